// Implementation of the reference-facing C++ interface (cgrt_host.h) on top of the C ABI, plus the small extern "C"
// surface that lets non-C++ hosts (tests, bench.py) reach the loader / presets / BMP writer.
#include "cgrt_host.h"
#include "cgrt_b200.h"
#include "cgrt_host_c.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

// ---- flatten Scene -> cgrt_scene_desc ---------------------------------------------------------------------------------------
namespace {
struct FlatScene {
    std::vector<int32_t> vcount, tcount;
    std::vector<float> vertices, materials, spheres;
    std::vector<uint32_t> triangles;
    cgrt_scene_desc desc{};
    void build(const Scene& s)
    {
        for (const Mesh& m : s.meshes) {
            vcount.push_back((int32_t)m.vertices.size());
            tcount.push_back((int32_t)m.triangles.size());
            for (const Vertex& v : m.vertices) {
                const float f[6] = {v.p.x, v.p.y, v.p.z, v.n.x, v.n.y, v.n.z};
                vertices.insert(vertices.end(), f, f + 6);
            }
            for (const Triangle& t : m.triangles) {
                triangles.push_back(t.x);
                triangles.push_back(t.y);
                triangles.push_back(t.z);
            }
            const Material& a = m.material;
            const float f[8] = {a.kd.x, a.kd.y, a.kd.z, a.ks.x, a.ks.y, a.ks.z, a.shininess, a.transparency};
            materials.insert(materials.end(), f, f + 8);
        }
        packSpheres(s, spheres);
        desc.n_meshes = (int32_t)s.meshes.size();
        desc.mesh_vertex_count = vcount.data();
        desc.mesh_triangle_count = tcount.data();
        desc.vertices = vertices.data();
        desc.triangles = triangles.data();
        desc.materials = materials.data();
        desc.n_spheres = (int32_t)s.spheres.size();
        desc.spheres = spheres.data();
    }
    static void packSpheres(const Scene& s, std::vector<float>& out)
    {
        out.clear();
        for (const Sphere& sp : s.spheres) {
            const Material& a = sp.material;
            const float f[12] = {sp.center.x, sp.center.y, sp.center.z, sp.radius, a.kd.x, a.kd.y, a.kd.z,
                                 a.ks.x,      a.ks.y,      a.ks.z,      a.shininess, a.transparency};
            out.insert(out.end(), f, f + 12);
        }
    }
};

[[noreturn]] void throwLast(const char* what)
{
    throw std::runtime_error(std::string(what) + ": " + cgrt_last_error());
}

cgrt_ray toRay(const Ray& r)
{
    cgrt_ray c;
    c.origin[0] = r.origin.x; c.origin[1] = r.origin.y; c.origin[2] = r.origin.z;
    c.t = r.t;
    c.direction[0] = r.direction.x; c.direction[1] = r.direction.y; c.direction[2] = r.direction.z;
    c.pad = 0.0f;
    return c;
}
int g_defaultDevice = 0;
} // namespace

// ---- BoundingVolumeHierarchy ----------------------------------------------------------------------------------------------------
BoundingVolumeHierarchy::BoundingVolumeHierarchy(Scene* pScene) : BoundingVolumeHierarchy(pScene, g_defaultDevice, 12) {}

BoundingVolumeHierarchy::BoundingVolumeHierarchy(Scene* pScene, int device, int maxDepth) : m_pScene(pScene), m_device(device)
{
    FlatScene flat;
    flat.build(*pScene);
    cgrt_scene_options opt;
    std::memset(&opt, 0, sizeof opt);
    opt.device = device;
    opt.bvh_max_depth = maxDepth;
    if (cgrt_scene_create(&flat.desc, &opt, &m_handle) != CGRT_OK) throwLast("BoundingVolumeHierarchy");
    m_sphereCache = flat.spheres;
    for (size_t m = 0; m < pScene->meshes.size(); m++) {
        m_materials.push_back(pScene->meshes[m].material);
        m_triToMesh.insert(m_triToMesh.end(), pScene->meshes[m].triangles.size(), (int32_t)m);
    }
}

BoundingVolumeHierarchy::~BoundingVolumeHierarchy()
{
    if (m_handle) cgrt_scene_destroy(m_handle);
}

BoundingVolumeHierarchy::BoundingVolumeHierarchy(BoundingVolumeHierarchy&& o) noexcept { *this = std::move(o); }
BoundingVolumeHierarchy& BoundingVolumeHierarchy::operator=(BoundingVolumeHierarchy&& o) noexcept
{
    if (this != &o) {
        if (m_handle) cgrt_scene_destroy(m_handle);
        m_pScene = o.m_pScene;
        m_handle = o.m_handle;
        m_device = o.m_device;
        m_materials = std::move(o.m_materials);
        m_triToMesh = std::move(o.m_triToMesh);
        m_sphereCache = std::move(o.m_sphereCache);
        o.m_handle = nullptr;
    }
    return *this;
}

int BoundingVolumeHierarchy::numLevels() const { return cgrt_bvh_num_levels(m_handle); }
void BoundingVolumeHierarchy::debugDraw(int) {}

std::vector<BoundingVolumeHierarchy::DebugNode> BoundingVolumeHierarchy::debugNodes(int level) const
{ // getNodesAtLevel, bvh.cpp:446-462: the nodes whose level equals `level` (a leaf above that level contributes nothing)
    const int n = cgrt_bvh_num_nodes(m_handle);
    std::vector<int32_t> meta((size_t)n * 5);
    std::vector<float> aabb((size_t)n * 6);
    std::vector<DebugNode> out;
    if (n == 0 || cgrt_bvh_export_nodes(m_handle, meta.data(), aabb.data()) != CGRT_OK) return out;
    for (int i = 0; i < n; i++) {
        if (meta[5 * i + 1] != level) continue;
        DebugNode d;
        d.AABB.lower = glm::vec3(aabb[6 * i], aabb[6 * i + 1], aabb[6 * i + 2]);
        d.AABB.upper = glm::vec3(aabb[6 * i + 3], aabb[6 * i + 4], aabb[6 * i + 5]);
        d.isLeaf = meta[5 * i] != 0;
        d.level = level;
        out.push_back(d);
    }
    return out;
}

void BoundingVolumeHierarchy::syncSpheres() const
{ // the reference reads m_pScene->spheres at query time (bvh.cpp:878): re-upload when they changed
    std::vector<float> now;
    FlatScene::packSpheres(*m_pScene, now);
    if (now.size() != m_sphereCache.size() || std::memcmp(now.data(), m_sphereCache.data(), now.size() * sizeof(float)) != 0) {
        if (cgrt_scene_set_spheres(m_handle, now.data(), (int32_t)(now.size() / 12)) != CGRT_OK) throwLast("set_spheres");
        m_sphereCache = now;
    }
}

bool BoundingVolumeHierarchy::intersect(Ray& ray, HitInfo& hitInfo) const
{
    syncSpheres();
    cgrt_ray r = toRay(ray);
    cgrt_hit h;
    if (cgrt_intersect_closest(m_handle, &r, 1, &h, nullptr) != CGRT_OK) throwLast("intersect");
    if (h.tri == -1) return false; // miss: ray.t and hitInfo untouched
    ray.t = h.t;
    hitInfo.normal = glm::vec3(h.normal[0], h.normal[1], h.normal[2]);
    int32_t materialTri = h.tri;
    if (h.tri <= -2) std::memcpy(&materialTri, &h.alpha, 4); // sphere hit: material of the last accepted triangle, if any
    if (materialTri >= 0) hitInfo.material = m_materials[m_triToMesh[materialTri]];
    return true;
}

// ---- free functions of ray_tracing.h: batch-of-one launches of the same device predicates ---------------------------------------
bool intersectRayWithPlane(const Plane& plane, Ray& ray)
{
    const float p[4] = {plane.normal.x, plane.normal.y, plane.normal.z, plane.D};
    cgrt_ray r = toRay(ray);
    uint8_t hit = 0;
    float t = ray.t;
    if (cgrt_ray_plane(g_defaultDevice, p, &r, 1, &hit, &t) != CGRT_OK) throwLast("intersectRayWithPlane");
    if (hit) ray.t = t;
    return hit != 0;
}

bool pointInTriangle(const glm::vec3& v0, const glm::vec3& v1, const glm::vec3& v2, const glm::vec3& n, const glm::vec3& p)
{
    const float in[15] = {v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z, n.x, n.y, n.z, p.x, p.y, p.z};
    uint8_t inside = 0;
    if (cgrt_point_in_triangle(g_defaultDevice, in, 1, &inside) != CGRT_OK) throwLast("pointInTriangle");
    return inside != 0;
}

Plane trianglePlane(const glm::vec3& v0, const glm::vec3& v1, const glm::vec3& v2)
{
    const float in[9] = {v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z};
    float out[4];
    if (cgrt_triangle_plane(g_defaultDevice, in, 1, out) != CGRT_OK) throwLast("trianglePlane");
    Plane pl;
    pl.normal = glm::vec3(out[0], out[1], out[2]);
    pl.D = out[3];
    return pl;
}

bool intersectRayWithTriangle(const glm::vec3& v0, const glm::vec3& v1, const glm::vec3& v2, Ray& ray, HitInfo& hitInfo,
                              const glm::vec3& n1, const glm::vec3& n2, const glm::vec3& n3)
{
    const float in[18] = {v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z,
                          n1.x, n1.y, n1.z, n2.x, n2.y, n2.z, n3.x, n3.y, n3.z};
    cgrt_ray r = toRay(ray);
    cgrt_hit h;
    if (cgrt_ray_triangle(g_defaultDevice, in, &r, 1, &h) != CGRT_OK) throwLast("intersectRayWithTriangle");
    if (h.tri != 1) return false;
    ray.t = h.t;
    hitInfo.normal = glm::vec3(h.normal[0], h.normal[1], h.normal[2]);
    return true;
}

bool intersectRayWithShape(const Sphere& sphere, Ray& ray, HitInfo& hitInfo)
{
    const float s[4] = {sphere.center.x, sphere.center.y, sphere.center.z, sphere.radius};
    cgrt_ray r = toRay(ray);
    float out[5];
    if (cgrt_ray_sphere(g_defaultDevice, s, &r, 1, out) != CGRT_OK) throwLast("intersectRayWithShape(Sphere)");
    int32_t hit;
    std::memcpy(&hit, &out[1], 4);
    if (!hit) return false;
    ray.t = out[0];
    hitInfo.normal = glm::vec3(out[2], out[3], out[4]); // material is NOT set, as in the reference (ray_tracing.cpp:154-157)
    return true;
}

bool intersectRayWithShape(const AxisAlignedBox& box, Ray& ray)
{
    const float b[6] = {box.lower.x, box.lower.y, box.lower.z, box.upper.x, box.upper.y, box.upper.z};
    cgrt_ray r = toRay(ray);
    uint8_t hit = 0;
    float t = ray.t;
    if (cgrt_ray_aabb(g_defaultDevice, b, &r, 1, &hit, &t) != CGRT_OK) throwLast("intersectRayWithShape(AABB)");
    if (hit) ray.t = t;
    return hit != 0;
}

bool intersectRayWithShape(const Mesh& mesh, Ray& ray, HitInfo& hitInfo)
{ // brute force over one mesh (ray_tracing.cpp:202-213): a throw-away one-mesh scene queried with the brute-force kernel
    Scene tmp;
    tmp.meshes.push_back(mesh);
    FlatScene flat;
    flat.build(tmp);
    cgrt_scene_options opt;
    std::memset(&opt, 0, sizeof opt);
    opt.device = g_defaultDevice;
    cgrt_scene* s = nullptr;
    if (cgrt_scene_create(&flat.desc, &opt, &s) != CGRT_OK) throwLast("intersectRayWithShape(Mesh)");
    cgrt_ray r = toRay(ray);
    cgrt_hit h;
    const int rc = cgrt_intersect_brute(s, &r, 1, &h);
    cgrt_scene_destroy(s);
    if (rc != CGRT_OK) throwLast("intersectRayWithShape(Mesh)");
    if (h.tri < 0) return false;
    ray.t = h.t;
    hitInfo.normal = glm::vec3(h.normal[0], h.normal[1], h.normal[2]); // no material write (ray_tracing.cpp:202-213)
    return true;
}

// ---- Trackball --------------------------------------------------------------------------------------------------------------------
Trackball::Trackball(float aspectRatio, float fovy, float distanceFromLookAt, float rotationX, float rotationY)
    : m_aspect(aspectRatio), m_fovy(fovy), m_distanceFromLookAt(distanceFromLookAt)
{
    m_rotationEulerAngles.x = rotationX; // framework/src/trackball.cpp:24-25
    m_rotationEulerAngles.y = rotationY;
}
void Trackball::setCamera(const glm::vec3 lookAt, const glm::vec3 rotations, const float dist)
{
    m_lookAt = lookAt;
    m_rotationEulerAngles = rotations;
    m_distanceFromLookAt = dist;
}
void Trackball::setLookAt(const glm::vec3 lookAt) { m_lookAt = lookAt; }

namespace {
cgrt_camera toCamera(const Trackball& c)
{
    cgrt_camera k;
    k.fovy = c.fovy();
    k.aspect = c.aspectRatio();
    k.dist = c.distanceFromLookAt();
    const glm::vec3 l = c.lookAt(), e = c.rotationEulerAngles();
    k.look_at[0] = l.x; k.look_at[1] = l.y; k.look_at[2] = l.z;
    k.euler[0] = e.x; k.euler[1] = e.y; k.euler[2] = e.z;
    return k;
}
} // namespace

Ray Trackball::generateRay(const glm::vec2& pixel) const
{ // Host evaluation of trackball.cpp:92-103 for single rays (the debug ray from the mouse, main.cpp:747-753); frames go
  // through cgrt_render, whose ray generation runs on the device from the same host-evaluated constants.
    const float halfH = std::tan(m_fovy / 2.0f);
    const float halfW = m_aspect * halfH;
    const glm::vec3 e = m_rotationEulerAngles;
    const float cx = std::cos(e.x * 0.5f), cy = std::cos(e.y * 0.5f), cz = std::cos(e.z * 0.5f);
    const float sx = std::sin(e.x * 0.5f), sy = std::sin(e.y * 0.5f), sz = std::sin(e.z * 0.5f);
    const float qw = cx * cy * cz + sx * sy * sz, qx = sx * cy * cz - cx * sy * sz, qy = cx * sy * cz + sx * cy * sz,
                qz = cx * cy * sz - sx * sy * cz;
    auto cross = [](const glm::vec3& a, const glm::vec3& b) {
        return glm::vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
    };
    auto rot = [&](const glm::vec3& v) {
        const glm::vec3 q(qx, qy, qz);
        const glm::vec3 uv = cross(q, v), uuv = cross(q, uv);
        return v + ((uv * qw) + uuv) * 2.0f;
    };
    glm::vec3 d(-pixel.x * halfW, pixel.y * halfH, 1.0f);
    const float inv = 1.0f / std::sqrt((d.x * d.x + d.y * d.y) + d.z * d.z);
    d = d * inv;
    Ray ray;
    ray.origin = m_lookAt + rot(glm::vec3(0.0f, 0.0f, -m_distanceFromLookAt));
    ray.direction = rot(d);
    ray.t = std::numeric_limits<float>::max();
    return ray;
}
glm::vec3 Trackball::position() const { return generateRay(glm::vec2(0.0f, 0.0f)).origin; }

// ---- Screen -----------------------------------------------------------------------------------------------------------------------
Screen::Screen(const glm::ivec2& resolution)
    : m_resolution(resolution), m_textureData(size_t(resolution.x) * size_t(resolution.y), glm::vec3(0.0f))
{
}
void Screen::clear(const glm::vec3& color) { std::fill(m_textureData.begin(), m_textureData.end(), color); }
void Screen::setPixel(int x, int y, const glm::vec3& color)
{
    const int i = (m_resolution.y - 1 - y) * m_resolution.x + x; // screen.cpp:34
    m_textureData[i] = color;
}
void Screen::writeBitmapToFile(const std::filesystem::path& filePath)
{
    cgrt_write_bmp(filePath.string().c_str(), data(), m_resolution.x, m_resolution.y);
}

// ---- renderRayTracing -------------------------------------------------------------------------------------------------------------
bool bloom = false, blur = false, antiAliasing = false; // src/main.cpp:33-35

RenderReport renderRayTracing(const Scene& scene, const Trackball& camera, const BoundingVolumeHierarchy& bvh, Screen& screen,
                              const RenderOptions& opt)
{
    std::vector<cgrt_point_light> lights;
    for (const PointLight& l : scene.pointLights) {
        cgrt_point_light c;
        c.position[0] = l.position.x; c.position[1] = l.position.y; c.position[2] = l.position.z;
        c.color[0] = l.color.x; c.color[1] = l.color.y; c.color[2] = l.color.z;
        lights.push_back(c);
    }
    if (cgrt_scene_set_lights(bvh.handle(), lights.data(), (int32_t)lights.size()) != CGRT_OK) throwLast("set_lights");
    {
        std::vector<float> sph; // Scene::sphericalLight, read live like the point lights
        for (const SphericalLight& l : scene.sphericalLight) {
            const float q[7] = {l.position.x, l.position.y, l.position.z, l.radius, l.color.x, l.color.y, l.color.z};
            sph.insert(sph.end(), q, q + 7);
        }
        if (cgrt_scene_set_spherical_lights(bvh.handle(), sph.data(), (int32_t)(sph.size() / 7), 1u) != CGRT_OK) throwLast("set_spherical_lights");
    }
    if (!scene.spheres.empty() || bvh.scene() == &scene) {
        // spheres are read live through the scene (bvh.cpp:878)
        std::vector<float> sp;
        FlatScene::packSpheres(scene, sp);
        if (cgrt_scene_set_spheres(bvh.handle(), sp.data(), (int32_t)(sp.size() / 12)) != CGRT_OK) throwLast("set_spheres");
    }
    cgrt_camera cam = toCamera(camera);
    cgrt_render_params p;
    std::memset(&p, 0, sizeof p);
    p.width = screen.resolution().x;
    p.height = screen.resolution().y;
    p.trace_limit = opt.traceLimit;
    p.rank = opt.rank;
    p.world = opt.world;
    cgrt_render_stats st;
    const int32_t effects = (antiAliasing ? CGRT_EFFECT_ANTIALIAS : 0) | (blur ? CGRT_EFFECT_MOTION_BLUR : 0) | (bloom ? CGRT_EFFECT_BLOOM : 0);
    if (effects && opt.world == 1) { // anti-aliasing / bloom / motion blur around the same renderer (main.cpp:663-687, :586-628, :318-584)
        if (cgrt_render_effects(bvh.handle(), &cam, &p, effects, screen.data(), &st) != CGRT_OK) throwLast("renderRayTracing");
    } else if (cgrt_render(bvh.handle(), &cam, &p, screen.data(), &st) != CGRT_OK) throwLast("renderRayTracing");
    RenderReport r;
    r.primary = st.primary; r.primaryHit = st.primary_hit; r.shadow = st.shadow; r.bounce = st.bounce;
    r.kernelLaunches = st.kernel_launches;
    r.deviceMs = st.device_ms;
    return r;
}

void renderRayTracing(const Scene& scene, const Trackball& camera, const BoundingVolumeHierarchy& bvh, Screen& screen)
{
    (void)renderRayTracing(scene, camera, bvh, screen, RenderOptions());
}

// ---- extern "C" extras for non-C++ hosts ---------------------------------------------------------------------------------------------
struct cgrt_host_scene {
    Scene scene;
    FlatScene flat;
    std::string name;
};

extern "C" {

void cgrt_host_set_default_device(int device) { g_defaultDevice = device; }

static int finishHostScene(cgrt_host_scene* hs, cgrt_host_scene** out)
{
    hs->flat.build(hs->scene);
    *out = hs;
    return CGRT_OK;
}

int cgrt_host_scene_load_preset(const char* preset, const char* data_dir, cgrt_host_scene** out)
{
    if (!preset || !data_dir || !out) return CGRT_ERR_INVALID;
    static const struct { const char* name; SceneType type; } table[] = {
        {"SingleTriangle", SingleTriangle}, {"Cube", Cube}, {"CornellBox", CornellBox},
        {"CornellBoxSphericalLight", CornellBoxSphericalLight}, {"Monkey", Monkey}, {"Dragon", Dragon},
        {"Spheres", Spheres}, {"Custom", Custom}};
    for (const auto& e : table) {
        if (std::strcmp(e.name, preset) != 0) continue;
        cgrt_host_scene* hs = new cgrt_host_scene();
        hs->name = preset;
        try {
            hs->scene = loadScene(e.type, data_dir);
        } catch (const std::exception&) {
            delete hs;
            return CGRT_ERR_INVALID;
        }
        return finishHostScene(hs, out);
    }
    return CGRT_ERR_INVALID;
}

int cgrt_host_scene_load_obj(const char* path, int normalize, cgrt_host_scene** out)
{
    if (!path || !out) return CGRT_ERR_INVALID;
    cgrt_host_scene* hs = new cgrt_host_scene();
    hs->name = path;
    try {
        hs->scene.meshes = loadMesh(path, normalize != 0);
    } catch (const std::exception&) {
        delete hs;
        return CGRT_ERR_INVALID;
    }
    return finishHostScene(hs, out);
}

int cgrt_host_scene_dragon_standin(int segments_u, int segments_v, cgrt_host_scene** out)
{
    if (!out || segments_u < 3 || segments_v < 3) return CGRT_ERR_INVALID;
    cgrt_host_scene* hs = new cgrt_host_scene();
    hs->name = "dragon-standin";
    Scene s;
    s.meshes = makeDragonStandIn(segments_u, segments_v);
    centerAndScaleToUnitMesh(s.meshes);                                            // as loadMesh(file, true) would
    s.pointLights.push_back(PointLight{glm::vec3(-1, 1, -1), glm::vec3(1)});       // Dragon preset light, src/scene.cpp:42-44
    hs->scene = std::move(s);
    return finishHostScene(hs, out);
}

void cgrt_host_scene_destroy(cgrt_host_scene* hs) { delete hs; }

int cgrt_host_scene_desc(const cgrt_host_scene* hs, cgrt_scene_desc* out)
{
    if (!hs || !out) return CGRT_ERR_INVALID;
    *out = hs->flat.desc;
    return CGRT_OK;
}

int64_t cgrt_host_scene_counts(const cgrt_host_scene* hs, int64_t* n_vertices, int64_t* n_triangles)
{
    if (!hs) return 0;
    if (n_vertices) *n_vertices = (int64_t)hs->flat.vertices.size() / 6;
    if (n_triangles) *n_triangles = (int64_t)hs->flat.triangles.size() / 3;
    return (int64_t)hs->scene.meshes.size();
}

int cgrt_host_scene_lights(const cgrt_host_scene* hs, cgrt_point_light* out, int cap)
{
    if (!hs) return 0;
    int n = 0;
    for (const PointLight& l : hs->scene.pointLights) {
        if (out && n < cap) {
            out[n].position[0] = l.position.x; out[n].position[1] = l.position.y; out[n].position[2] = l.position.z;
            out[n].color[0] = l.color.x; out[n].color[1] = l.color.y; out[n].color[2] = l.color.z;
        }
        n++;
    }
    return n;
}

// Screen::writeBitmapToFile (src/screen.cpp:38-49): clamp to [0,1], * 255, truncate to 8 bit, alpha 255, rows top to bottom
// as stored in the Screen layout. stb_image_write's stbi_write_bmp(.., comp = 4, ..) emits a 32-bit BMP (BITMAPV4HEADER
// with channel masks); a plain 32-bit BI_RGB file with the same pixels is written here.
int cgrt_write_bmp(const char* path, const float* rgb, int width, int height)
{
    if (!path || !rgb || width <= 0 || height <= 0) return CGRT_ERR_INVALID;
    FILE* f = std::fopen(path, "wb");
    if (!f) return CGRT_ERR_INVALID;
    const uint32_t rowBytes = (uint32_t)width * 4, dataBytes = rowBytes * (uint32_t)height, off = 14 + 40;
    unsigned char hdr[54];
    std::memset(hdr, 0, sizeof hdr);
    auto put32 = [&](int at, uint32_t v) { hdr[at] = v & 255; hdr[at + 1] = (v >> 8) & 255; hdr[at + 2] = (v >> 16) & 255; hdr[at + 3] = (v >> 24) & 255; };
    hdr[0] = 'B'; hdr[1] = 'M';
    put32(2, off + dataBytes);
    put32(10, off);
    put32(14, 40);
    put32(18, (uint32_t)width);
    put32(22, (uint32_t)height); // positive height = bottom-up rows
    hdr[26] = 1;
    hdr[28] = 32;
    put32(34, dataBytes);
    std::fwrite(hdr, 1, sizeof hdr, f);
    std::vector<unsigned char> row(rowBytes);
    for (int y = height - 1; y >= 0; y--) { // file rows bottom-up; Screen row 0 is the top of the image
        for (int x = 0; x < width; x++) {
            unsigned char px[3];
            for (int k = 0; k < 3; k++) {
                float v = rgb[3 * ((size_t)y * width + x) + k];
                v = v < 0.0f ? 0.0f : v;
                v = 1.0f < v ? 1.0f : v;
                px[k] = (unsigned char)(v * 255.0f);
            }
            row[4 * x + 0] = px[2]; row[4 * x + 1] = px[1]; row[4 * x + 2] = px[0]; row[4 * x + 3] = 255;
        }
        std::fwrite(row.data(), 1, rowBytes, f);
    }
    std::fclose(f);
    return CGRT_OK;
}

} // extern "C"
