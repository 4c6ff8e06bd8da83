// Test-infrastructure shim (NOT product code, NOT glm): the minimal subset of the glm 0.9.9.8
// vec3/uvec3 interface that /root/reference/src/{ray_tracing,bounding_volume_hierarchy}.cpp use,
// so that those two reference TUs compile verbatim without the (un-fetchable) glm dependency.
// Arithmetic follows glm 0.9.9.8's scalar code paths (component-wise, no reciprocal tricks);
// glm itself is pinned by the reference at framework/cmake/download_framework_packages.cmake:19-22
// and is absent from /root/reference => fidelity to real glm is "parity unpinned" (see DESIGN.md).
#pragma once
#include <cstddef>

namespace glm {

struct vec3 {
    union { float x; float r; };
    union { float y; float g; };
    union { float z; float b; };

    constexpr vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    constexpr explicit vec3(float s) : x(s), y(s), z(s) {}
    constexpr vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
    template <typename A, typename B, typename C>
    constexpr vec3(A a, B b_, C c) : x(static_cast<float>(a)), y(static_cast<float>(b_)), z(static_cast<float>(c)) {}
    constexpr vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
    vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }

    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }

    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

struct uvec3 {
    unsigned x, y, z;
    constexpr uvec3() : x(0), y(0), z(0) {}
    constexpr uvec3(unsigned a, unsigned b, unsigned c) : x(a), y(b), z(c) {}
    unsigned& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const unsigned& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

} // namespace glm
