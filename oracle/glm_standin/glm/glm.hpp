// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for the subset of the
// header-only GLM maths library that the reference's hot-path translation units
// use (cuda_rasterizer/{forward,backward,rasterizer_impl}.cu).  GLM is an
// un-vendored third-party dependency of the reference (no glm.hpp exists in
// /root/reference nor in this image), so the reference CUDA cannot be compiled
// as the GPU oracle without it.  Semantics follow GLM's published conventions:
//   * mat3 is COLUMN-major: m[c][r]; the 9-scalar constructor fills column 0
//     first, then column 1, then column 2.
//   * mat3*mat3 element (c,r) = a[0][r]*b[c][0] + a[1][r]*b[c][1] + a[2][r]*b[c][2],
//     summed left to right (so nvcc's FMA contraction has the same shape).
//   * dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z summed left to right; length = sqrt(dot).
// Nothing in the product path (gsdf_slam_b200/) includes this file.
#pragma once
#include <cfloat>
#include <cmath>
#include <cuda_runtime.h>

#define GLM_STANDIN_FN __host__ __device__ __forceinline__

namespace glm {

struct vec2 {
  float x, y;
  GLM_STANDIN_FN vec2() : x(0), y(0) {}
  template <typename A, typename B>
  GLM_STANDIN_FN vec2(A a, B b) : x(float(a)), y(float(b)) {}
};

struct vec3 {
  float x, y, z;
  GLM_STANDIN_FN vec3() : x(0), y(0), z(0) {}
  GLM_STANDIN_FN explicit vec3(float s) : x(s), y(s), z(s) {}
  template <typename A, typename B, typename C>
  GLM_STANDIN_FN vec3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
  GLM_STANDIN_FN float& operator[](int i) { return (&x)[i]; }
  GLM_STANDIN_FN const float& operator[](int i) const { return (&x)[i]; }
  GLM_STANDIN_FN vec3& operator+=(const vec3& o) {
    x += o.x; y += o.y; z += o.z; return *this;
  }
  GLM_STANDIN_FN vec3& operator+=(float s) {
    x += s; y += s; z += s; return *this;
  }
  GLM_STANDIN_FN vec3& operator*=(float s) {
    x *= s; y *= s; z *= s; return *this;
  }
  GLM_STANDIN_FN vec3& operator=(const vec3&) = default;
  GLM_STANDIN_FN vec3(const vec3&) = default;
};

struct vec4 {
  float x, y, z, w;
  GLM_STANDIN_FN vec4() : x(0), y(0), z(0), w(0) {}
  template <typename A, typename B, typename C, typename D>
  GLM_STANDIN_FN vec4(A a, B b, C c, D d)
      : x(float(a)), y(float(b)), z(float(c)), w(float(d)) {}
};

GLM_STANDIN_FN vec3 operator+(const vec3& a, const vec3& b) {
  return vec3(a.x + b.x, a.y + b.y, a.z + b.z);
}
GLM_STANDIN_FN vec3 operator-(const vec3& a, const vec3& b) {
  return vec3(a.x - b.x, a.y - b.y, a.z - b.z);
}
GLM_STANDIN_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_STANDIN_FN vec3 operator*(float s, const vec3& a) {
  return vec3(s * a.x, s * a.y, s * a.z);
}
GLM_STANDIN_FN vec3 operator*(const vec3& a, float s) {
  return vec3(a.x * s, a.y * s, a.z * s);
}
GLM_STANDIN_FN vec3 operator/(const vec3& a, float s) {
  return vec3(a.x / s, a.y / s, a.z / s);
}
GLM_STANDIN_FN float dot(const vec3& a, const vec3& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
GLM_STANDIN_FN float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLM_STANDIN_FN float length(const vec4& a) {
  return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
}
GLM_STANDIN_FN vec3 max(const vec3& a, float s) {
  return vec3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s));
}

struct mat3 {
  vec3 c[3];  // columns
  GLM_STANDIN_FN mat3() {
    c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1);
  }
  GLM_STANDIN_FN explicit mat3(float s) {
    c[0] = vec3(s, 0, 0); c[1] = vec3(0, s, 0); c[2] = vec3(0, 0, s);
  }
  template <typename X1, typename Y1, typename Z1, typename X2, typename Y2,
            typename Z2, typename X3, typename Y3, typename Z3>
  GLM_STANDIN_FN mat3(X1 x1, Y1 y1, Z1 z1, X2 x2, Y2 y2, Z2 z2, X3 x3, Y3 y3,
                      Z3 z3) {
    c[0] = vec3(x1, y1, z1); c[1] = vec3(x2, y2, z2); c[2] = vec3(x3, y3, z3);
  }
  GLM_STANDIN_FN vec3& operator[](int i) { return c[i]; }
  GLM_STANDIN_FN const vec3& operator[](int i) const { return c[i]; }
};

GLM_STANDIN_FN mat3 transpose(const mat3& m) {
  return mat3(m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2],
              m[1][2], m[2][2]);
}

GLM_STANDIN_FN mat3 operator*(const mat3& a, const mat3& b) {
  mat3 r;
  for (int col = 0; col < 3; ++col)
    for (int row = 0; row < 3; ++row)
      r[col][row] =
          a[0][row] * b[col][0] + a[1][row] * b[col][1] + a[2][row] * b[col][2];
  return r;
}
GLM_STANDIN_FN mat3 operator*(float s, const mat3& m) {
  mat3 r;
  r[0] = m[0] * s; r[1] = m[1] * s; r[2] = m[2] * s;
  return r;
}
GLM_STANDIN_FN mat3 operator*(const mat3& m, float s) { return s * m; }
GLM_STANDIN_FN vec3 operator*(const mat3& m, const vec3& v) {
  return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
              m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
              m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}

}  // namespace glm
