// Stand-in for <glm/vec3.hpp> used ONLY when the real glm (0.9.9.8, the reference's dependency) is not on the include
// path: the drop-in host layer keeps the reference's signatures, which are written in terms of glm::vec3 / glm::uvec3.
// A maintainer integrating into the reference tree builds against the real glm and drops this directory.
#pragma once
namespace glm {
struct vec3 {
    float x, y, z;
    constexpr vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    constexpr explicit vec3(float s) : x(s), y(s), z(s) {}
    constexpr vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct uvec3 {
    unsigned x, y, z;
    constexpr uvec3() : x(0), y(0), z(0) {}
    constexpr uvec3(unsigned a, unsigned b, unsigned c) : x(a), y(b), z(c) {}
    unsigned& operator[](int i) { return (&x)[i]; }
    const unsigned& operator[](int i) const { return (&x)[i]; }
};
struct ivec2 {
    int x, y;
    constexpr ivec2() : x(0), y(0) {}
    constexpr ivec2(int a, int b) : x(a), y(b) {}
};
struct vec2 {
    float x, y;
    constexpr vec2() : x(0.0f), y(0.0f) {}
    constexpr vec2(float a, float b) : x(a), y(b) {}
};
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
constexpr float radians(float degrees) { return degrees * 0.01745329251994329576923690768489f; }
} // namespace glm
