// TEST INFRASTRUCTURE ONLY -- no-op stand-in for Dear ImGui (see oracle/shim_app/README.md).
#pragma once
struct ImVec2 {
  float x = 0.f, y = 0.f;
  ImVec2() = default;
  ImVec2(float a, float b) : x(a), y(b) {}
};
struct ImVec4 {
  float x, y, z, w;
  ImVec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
struct ImGuiIO {
  ImVec2 DisplaySize{1280.f, 720.f};
  float Framerate = 60.f;
};
namespace ImGui {
inline ImGuiIO &GetIO() {
  static ImGuiIO io;
  return io;
}
inline bool BeginMainMenuBar() { return false; }
inline void EndMainMenuBar() {}
inline bool BeginMenu(const char *, bool = true) { return false; }
inline void EndMenu() {}
inline bool MenuItem(const char *, const char * = nullptr, bool = false, bool = true) { return false; }
inline void OpenPopup(const char *, int = 0) {}
inline bool Begin(const char *, bool * = nullptr, int = 0) { return true; }
inline void End() {}
inline void Text(const char *, ...) {}
inline void SameLine(float = 0.f, float = -1.f) {}
inline bool Checkbox(const char *, bool *) { return false; }
inline bool Button(const char *, const ImVec2 & = ImVec2()) { return false; }
inline bool SliderFloat(const char *, float *, float, float, const char * = "%.3f", int = 0) { return false; }
inline bool InputDouble(const char *, double *, double = 0., double = 0., const char * = "%.6f", int = 0) { return false; }
} // namespace ImGui
