// TEST INFRASTRUCTURE ONLY (oracle build shim) -- not part of the product path.
// pressure_solver.cpp:2 includes draw_2dbuf.hpp -> texture.hpp, whose inline
// constructor names a handful of OpenGL entry points.  Nothing on the solver
// path ever calls them; these no-op stand-ins only let the TU compile.
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef char GLchar;
#define GL_NO_ERROR 0
#define GL_TEXTURE_2D 0x0DE1
#define GL_SRGB8_ALPHA8 0x8C43
#define GL_RGBA 0x1908
#define GL_UNSIGNED_BYTE 0x1401
inline GLenum glGetError() { return GL_NO_ERROR; }
inline void glGenTextures(GLsizei, GLuint *) {}
inline void glBindTexture(GLenum, GLuint) {}
inline void glTexStorage2D(GLenum, GLsizei, GLenum, GLsizei, GLsizei) {}
inline void glTexSubImage2D(GLenum, GLint, GLint, GLint, GLsizei, GLsizei, GLenum, GLenum, const void *) {}
inline void glGenerateMipmap(GLenum) {}
