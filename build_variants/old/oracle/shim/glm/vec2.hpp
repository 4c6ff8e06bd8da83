// Test-infrastructure shim: glm::vec2 placeholder (mesh.h includes it; the hot-path TUs never use it).
#pragma once
namespace glm { struct vec2 { float x, y; }; }
