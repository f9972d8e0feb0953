// TEST INFRASTRUCTURE ONLY -- fixed-function OpenGL names used by app.cpp / texture.hpp / spec-cache.cpp
// as no-ops, except glTexImage1D, which records the texels it is given (oracle/ref_app.cpp reads them
// back to pin the colour ramp of SpecCache::populateTex, spec-cache.cpp:77-96).
#pragma once
#include <cstring>
#include <vector>
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef unsigned int GLbitfield;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
typedef double GLdouble;
enum : GLenum {
  GL_BLEND = 0x0BE2, GL_COLOR_BUFFER_BIT = 0x4000, GL_LINES = 1, GL_LINE_STRIP = 3, GL_QUADS = 7, GL_NEAREST = 0x2600,
  GL_SRC_ALPHA = 0x0302, GL_ONE_MINUS_SRC_ALPHA = 0x0303, GL_PROJECTION = 0x1701, GL_RGB = 0x1907,
  GL_TEXTURE_1D = 0x0DE0, GL_TEXTURE_MAG_FILTER = 0x2800, GL_TEXTURE_MIN_FILTER = 0x2801, GL_UNSIGNED_BYTE = 0x1401
};
struct MlxGlCapture {  // last glTexImage1D upload
  std::vector<unsigned char> rgb;
  int width = 0;
  int uploads = 0;
};
inline MlxGlCapture &mlx_gl_capture() {
  static MlxGlCapture c;
  return c;
}
inline void glGenTextures(GLsizei n, GLuint *t) {
  static GLuint next = 1;
  for (GLsizei i = 0; i < n; ++i) t[i] = next++;
}
inline void glDeleteTextures(GLsizei, const GLuint *) {}
inline void glBindTexture(GLenum, GLuint) {}
inline void glTexParameteri(GLenum, GLenum, GLint) {}
inline void glTexImage1D(GLenum, GLint, GLint, GLsizei width, GLint, GLenum, GLenum, const void *data) {
  MlxGlCapture &c = mlx_gl_capture();
  c.width = width;
  c.rgb.resize((size_t)width * 3);
  if (width > 0) std::memcpy(c.rgb.data(), data, (size_t)width * 3);
  ++c.uploads;
}
inline void glBegin(GLenum) {}
inline void glEnd() {}
inline void glEnable(GLenum) {}
inline void glDisable(GLenum) {}
inline void glBlendFunc(GLenum, GLenum) {}
inline void glClear(GLbitfield) {}
inline void glClearColor(GLfloat, GLfloat, GLfloat, GLfloat) {}
inline void glColor3f(GLfloat, GLfloat, GLfloat) {}
inline void glColor4f(GLfloat, GLfloat, GLfloat, GLfloat) {}
inline void glLoadIdentity() {}
inline void glMatrixMode(GLenum) {}
inline void glOrtho(GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble) {}
inline void glTexCoord1f(GLfloat) {}
inline void glVertex2f(GLfloat, GLfloat) {}
inline void glViewport(GLint, GLint, GLsizei, GLsizei) {}
