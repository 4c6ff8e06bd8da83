// Host-side mirror of the reference's C++ interface for the hot path. Same type names, same signatures, same argument
// meaning and error behaviour as the reference headers, implemented on top of the C ABI (include/cgrt_b200.h); every
// query is answered by the sm_100a kernels. A maintainer swaps these declarations in for the reference headers
// (see INTEGRATION.md); nothing here needs OpenGL / GLFW / ImGui, so it also builds headless.
//
//   reference header                      what is mirrored here
//   framework/include/ray.h:9-13          Ray
//   src/mesh.h:12-35                      Vertex, Material, Triangle, Mesh, loadMesh
//   src/scene.h:12-63                     SceneType, Plane, AxisAlignedBox, Sphere, PointLight, SphericalLight, Scene, loadScene
//   src/ray_tracing.h:4-20                HitInfo, intersectRayWithPlane/Triangle/Shape, pointInTriangle, trianglePlane
//   src/bounding_volume_hierarchy.h:15-57 BoundingVolumeHierarchy (ctor, intersect, numLevels, debugDraw -> debugNodes)
//   framework/include/trackball.h:16-56   Trackball (camera state + generateRay; mouse/GL parts are out of scope)
//   src/screen.h:12-27                    Screen (float RGB framebuffer, setPixel, writeBitmapToFile; the GL blit is out of scope)
//   src/main.cpp:648                      renderRayTracing(const Scene&, const Trackball&, const BoundingVolumeHierarchy&, Screen&)
#pragma once
#include <glm/vec2.hpp>
#include <glm/vec3.hpp>

#include <cstdint>
#include <filesystem>
#include <limits>
#include <string>
#include <vector>

struct cgrt_scene;

// ---- framework/include/ray.h ------------------------------------------------------------------------------------------------
struct Ray {
    glm::vec3 origin{0.0f};
    glm::vec3 direction{0.0f, 0.0f, -1.0f};
    float t{std::numeric_limits<float>::max()};
};

// ---- src/mesh.h -------------------------------------------------------------------------------------------------------------
struct Vertex {
    glm::vec3 p; // position
    glm::vec3 n; // normal
};
struct Material {
    glm::vec3 kd;
    glm::vec3 ks{0.0f};
    float shininess{1.0f};
    float transparency{1.0f};
};
using Triangle = glm::uvec3;
struct Mesh {
    std::vector<Vertex> vertices;
    std::vector<Triangle> triangles;
    Material material;
};
// Throws std::exception when the file is missing or cannot be imported (src/mesh.cpp:60-71).
[[nodiscard]] std::vector<Mesh> loadMesh(const std::filesystem::path& file, bool normalize = false);
// src/mesh.cpp:143-166 (file-static in the reference; exposed because the stand-in mesh needs the same normalisation)
void centerAndScaleToUnitMesh(std::vector<Mesh>& meshes);

// ---- src/scene.h ------------------------------------------------------------------------------------------------------------
enum SceneType { SingleTriangle, Cube, CornellBox, CornellBoxSphericalLight, Monkey, Dragon, Spheres, Custom };
struct Plane {
    float D = 0.0f;
    glm::vec3 normal{0.0f, 1.0f, 0.0f};
};
struct AxisAlignedBox {
    glm::vec3 lower{0.0f};
    glm::vec3 upper{1.0f};
};
struct Sphere {
    glm::vec3 center{0.0f};
    float radius = 1.0f;
    Material material;
};
struct PointLight {
    glm::vec3 position;
    glm::vec3 color;
};
struct SphericalLight {
    glm::vec3 position;
    float radius;
    glm::vec3 color;
};
struct Scene {
    std::vector<Mesh> meshes;
    std::vector<Sphere> spheres;
    std::vector<PointLight> pointLights;
    std::vector<SphericalLight> sphericalLight; // soft shadows (main.cpp:168-218): 200 sample rays per hit and light, read live at render time
};
Scene loadScene(SceneType type, const std::filesystem::path& dataDir);

// Procedural stand-in for data/dragon.obj, which is absent from the reference checkout (.MISSING_LARGE_BLOBS:1):
// a closed (2,3) torus-knot tube with 87 040 triangles (the report quotes 87 K for the dragon) and smooth normals,
// one Mesh, per-corner vertices like the loader produces. Deterministic.
std::vector<Mesh> makeDragonStandIn(int segmentsU = 340, int segmentsV = 128);

// ---- src/ray_tracing.h ------------------------------------------------------------------------------------------------------
struct HitInfo {
    glm::vec3 normal;
    Material material;
};
bool intersectRayWithPlane(const Plane& plane, Ray& ray);
bool pointInTriangle(const glm::vec3& v0, const glm::vec3& v1, const glm::vec3& v2, const glm::vec3& n, const glm::vec3& p);
Plane trianglePlane(const glm::vec3& v0, const glm::vec3& v1, const glm::vec3& v2);
bool intersectRayWithTriangle(const glm::vec3& v0, const glm::vec3& v1, const glm::vec3& v2, Ray& ray, HitInfo& hitInfo,
                              const glm::vec3& n1, const glm::vec3& n2, const glm::vec3& n3);
bool intersectRayWithShape(const Sphere& sphere, Ray& ray, HitInfo& hitInfo);
bool intersectRayWithShape(const AxisAlignedBox& box, Ray& ray);
bool intersectRayWithShape(const Mesh& mesh, Ray& ray, HitInfo& hitInfo);

// ---- src/bounding_volume_hierarchy.h ------------------------------------------------------------------------------------------
class BoundingVolumeHierarchy {
public:
    struct DebugNode { // what debugDraw(level) would hand to drawAABB (bvh.cpp:469-525): box + leaf flag
        AxisAlignedBox AABB;
        bool isLeaf;
        int level;
    };

    // Flattens pScene->meshes, builds the BVH with the reference split rule and uploads it (cgrt_scene_create).
    // Keeps the raw non-owning Scene* like the reference (bvh.h:19): lights and spheres are read through it at query time.
    // Throws std::runtime_error when no CUDA device is usable (there is no CPU fallback).
    BoundingVolumeHierarchy(Scene* pScene);
    BoundingVolumeHierarchy(Scene* pScene, int device, int maxDepth);
    ~BoundingVolumeHierarchy();
    BoundingVolumeHierarchy(const BoundingVolumeHierarchy&) = delete;
    BoundingVolumeHierarchy& operator=(const BoundingVolumeHierarchy&) = delete;
    BoundingVolumeHierarchy(BoundingVolumeHierarchy&& o) noexcept;
    BoundingVolumeHierarchy& operator=(BoundingVolumeHierarchy&& o) noexcept;

    // debugDraw(level) needs a GL context in the reference; here it returns the boxes it would draw.
    std::vector<DebugNode> debugNodes(int level) const;
    void debugDraw(int level); // headless: no-op
    int numLevels() const;

    // Return true if something is hit; only hits closer than ray.t count; on a hit ray.t, hitInfo.normal and
    // hitInfo.material are updated, on a miss they are left untouched (bvh.cpp:850-881). One-ray launch of the batch kernel.
    bool intersect(Ray& ray, HitInfo& hitInfo) const;

    // additions (not in the reference): batch form and access for the renderer
    cgrt_scene* handle() const { return m_handle; }
    const Scene* scene() const { return m_pScene; }
    int device() const { return m_device; }

private:
    Scene* m_pScene = nullptr;
    cgrt_scene* m_handle = nullptr;
    int m_device = 0;
    std::vector<Material> m_materials;     // per mesh, copied at construction like the reference copies the meshes (bvh.cpp:50)
    std::vector<int32_t> m_triToMesh;      // global triangle id -> mesh
    void syncSpheres() const;
    mutable std::vector<float> m_sphereCache;
};

// ---- framework/include/trackball.h (camera state + ray generation only) --------------------------------------------------------
class Trackball {
public:
    // The reference takes a Window* only to read aspectRatio(); headless, the aspect is given directly.
    Trackball(float aspectRatio, float fovy, float distanceFromLookAt = 4.0f, float rotationX = 0.0f, float rotationY = 0.0f);
    void setCamera(const glm::vec3 lookAt, const glm::vec3 rotations, const float dist);
    void setLookAt(const glm::vec3 lookAt);
    void setAspectRatio(float a) { m_aspect = a; }
    [[nodiscard]] glm::vec3 position() const;
    [[nodiscard]] glm::vec3 lookAt() const { return m_lookAt; }
    [[nodiscard]] Ray generateRay(const glm::vec2& pixel) const;
    // additions: const accessors for the private state the renderer needs bit-exactly (SURVEY.md §7 "hard parts")
    float fovy() const { return m_fovy; }
    float aspectRatio() const { return m_aspect; }
    float distanceFromLookAt() const { return m_distanceFromLookAt; }
    glm::vec3 rotationEulerAngles() const { return m_rotationEulerAngles; }

private:
    float m_aspect;
    float m_fovy;
    glm::vec3 m_lookAt{0.0f};
    float m_distanceFromLookAt;
    glm::vec3 m_rotationEulerAngles{0};
};

// ---- src/screen.h (framebuffer + BMP writer; GL texture blit out of scope) ------------------------------------------------------
class Screen {
public:
    Screen(const glm::ivec2& resolution);
    void clear(const glm::vec3& color);
    void setPixel(int x, int y, const glm::vec3& color);
    void writeBitmapToFile(const std::filesystem::path& filePath);
    void draw() {} // headless: nothing to blit to
    // additions: bulk hand-off of a whole frame already in Screen layout (row H-1-y) and read access
    glm::ivec2 resolution() const { return m_resolution; }
    float* data() { return &m_textureData[0].x; }
    const float* data() const { return &m_textureData[0].x; }

private:
    glm::ivec2 m_resolution;
    std::vector<glm::vec3> m_textureData;
};

// ---- src/main.cpp:648 -----------------------------------------------------------------------------------------------------------
// Renders screen.resolution() pixels (the reference hard-codes 800x800, main.cpp:29) with the given recursion limit
// (reference literal 2, main.cpp:267). Point lights are read from `scene` at call time, as the reference does.
// the reference's UI toggles (src/main.cpp:33-35, ImGui check boxes :878-882); renderRayTracing reads them like the reference
// does: anti-aliasing (:663-687), bloom (:586-628, the in-place 21 x 21 recurrence, run as a wavefront on the device) and
// motion blur (:318-584). bloom together with antiAliasing is refused (the reference's combination reads an uninitialised sum).
extern bool bloom, blur, antiAliasing;

struct RenderOptions {
    int traceLimit = 2;
    int rank = 0, world = 1; // interleaved tile partition; world > 1 renders only this rank's tiles into `screen`
};
struct RenderReport {
    uint64_t primary = 0, primaryHit = 0, shadow = 0, bounce = 0, kernelLaunches = 0;
    float deviceMs = 0.0f;
};
void renderRayTracing(const Scene& scene, const Trackball& camera, const BoundingVolumeHierarchy& bvh, Screen& screen);
RenderReport renderRayTracing(const Scene& scene, const Trackball& camera, const BoundingVolumeHierarchy& bvh, Screen& screen,
                              const RenderOptions& opt);
