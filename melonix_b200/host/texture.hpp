// melonix_b200/host/texture.hpp -- RAII GL texture name used by SpecCache (role of reference
// texture.hpp:10-36).  With MELONIX_HEADLESS the GL entry points come from gl_headless.h.
#pragma once
#ifdef MELONIX_HEADLESS
#include "gl_headless.h"
#else
#include <imgui/imgui.h>
#if defined(IMGUI_IMPL_OPENGL_ES2)
#include <SDL_opengles2.h>
#else
#include <SDL_opengl.h>
#endif
#endif
#include <utility>

class Texture
{
public:
  Texture() { glGenTextures(1, &name); }
  ~Texture()
  {
    if (name)
      glDeleteTextures(1, &name);
  }
  Texture(const Texture &) = delete;
  Texture &operator=(const Texture &) = delete;
  Texture(Texture &&o) noexcept : name(std::exchange(o.name, 0)) {}
  Texture &operator=(Texture &&o) noexcept
  {
    if (this != &o)
    {
      if (name)
        glDeleteTextures(1, &name);
      name = std::exchange(o.name, 0);
    }
    return *this;
  }
  auto get() const -> GLuint { return name; }
  operator GLuint() const { return name; }

private:
  GLuint name = 0;
};
