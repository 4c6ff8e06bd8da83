// Test-infrastructure shim: glm 0.9.9.8 geometric functions restated (scalar paths).
//   dot       : tmp = a*b; (tmp.x + tmp.y) + tmp.z           (detail/func_geometric.inl compute_dot<vec3>)
//   cross     : (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   length    : sqrt(dot(v,v))
//   normalize : v * inversesqrt(dot(v,v)),  inversesqrt(x) = 1/sqrt(x)
//   reflect   : I - N * dot(N,I) * 2
#pragma once
#include "vec3.hpp"
#include <cmath>

namespace glm {
inline float dot(const vec3& a, const vec3& b) { vec3 t(a * b); return t.x + t.y + t.z; }
inline vec3 cross(const vec3& x, const vec3& y)
{
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 reflect(const vec3& I, const vec3& N) { return I - N * dot(N, I) * 2.0f; }
} // namespace glm
