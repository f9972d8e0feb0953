// melonix_b200/host/gl_headless.h -- stand-in for <SDL_opengl.h> when MELONIX_HEADLESS is defined
// (tests and the GPU box have no GL context).  Records 1-D texture uploads in memory so that
// SpecCache can be exercised end to end.
#pragma once
#include <map>
#include <vector>

using GLuint = unsigned int;
using GLenum = unsigned int;
using GLint = int;
using GLsizei = int;
constexpr GLenum GL_TEXTURE_1D = 0x0DE0, GL_TEXTURE_MAG_FILTER = 0x2800, GL_TEXTURE_MIN_FILTER = 0x2801;
constexpr GLenum GL_NEAREST = 0x2600, GL_RGB = 0x1907, GL_UNSIGNED_BYTE = 0x1401;

namespace gl_headless
{
struct State
{
  GLuint next = 1, bound = 0;
  std::map<GLuint, std::vector<unsigned char>> tex; // name -> RGB bytes of the last upload
};
inline auto state() -> State &
{
  static State s;
  return s;
}
} // namespace gl_headless

inline void glGenTextures(GLsizei n, GLuint *out)
{
  for (GLsizei i = 0; i < n; ++i)
    out[i] = gl_headless::state().next++;
}
inline void glDeleteTextures(GLsizei n, const GLuint *names)
{
  for (GLsizei i = 0; i < n; ++i)
    gl_headless::state().tex.erase(names[i]);
}
inline void glBindTexture(GLenum, GLuint name) { gl_headless::state().bound = name; }
inline void glTexParameteri(GLenum, GLenum, GLint) {}
inline void glTexImage1D(GLenum, GLint, GLint, GLsizei width, GLint, GLenum, GLenum, const void *data)
{
  auto &s = gl_headless::state();
  const auto *p = static_cast<const unsigned char *>(data);
  s.tex[s.bound].assign(p, p + static_cast<size_t>(width) * 3);
}
