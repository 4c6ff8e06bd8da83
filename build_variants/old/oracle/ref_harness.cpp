// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// `--impl reference` legs may load the library built from this file (oracle/_ref/libcgrt_ref.so).
//
// C-ABI harness around the reference's own hot-path translation units, compiled VERBATIM from
// /root/reference/src/ray_tracing.cpp and /root/reference/src/bounding_volume_hierarchy.cpp (the latter with the
// documented two-`return` fix at :629/:632 applied on the fly by oracle/Makefile; nothing is copied into this repo).
// What is verbatim reference code when called through this harness:
//   * BoundingVolumeHierarchy::BoundingVolumeHierarchy(Scene*)          bounding_volume_hierarchy.cpp:42-76  (mode 0)
//   * BoundingVolumeHierarchy::intersect / intersectDataStructure / ... bounding_volume_hierarchy.cpp:535-881
//   * intersectRayWith{Plane,Triangle,Shape(AABB|Sphere|Mesh)}, pointInTriangle, trianglePlane, area   ray_tracing.cpp
// What is RESTATED here because main.cpp / trackball.cpp / screen.cpp cannot be compiled headless (ImGui/GLFW/GL):
//   * shading recursion  main.cpp:61-135 (specular/diffuse/pointInShadow), :220-232 (point-light loop), :241-310
//   * camera             framework/src/trackball.cpp:70-73, 92-103  (+ glm quat(euler), quat*vec3)
//   * pixel loop         main.cpp:653-697 (non-AA, non-bloom branch), Screen::setPixel screen.cpp:30-36
//   * a range-based BVH builder that fills the REFERENCE's own `Node` structs (mode 1) so that the verbatim traversal
//     can run on scenes where the reference constructor's O(nodes x vertices) copies are infeasible
//     (bounding_volume_hierarchy.cpp:205-206); it calls the reference's own sortTrianglesByCentres /
//     sortMeshesByCentres (bvh.cpp:88-134) and is validated node-for-node against mode 0 in tests/.
// Triangle ids / test counters are obtained WITHOUT touching reference code via -Wl,--wrap on the cross-TU calls.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <array>
#include <queue>
#include <vector>
#include <filesystem>
#include <optional>
#include <sstream>
#ifdef _OPENMP
#include <omp.h>
#endif

#define private public // test-only: read BoundingVolumeHierarchy::nodes / m_pScene (bounding_volume_hierarchy.h:18-23)
#include "bounding_volume_hierarchy.h"
#undef private
#include "draw.h"
#include <glm/geometric.hpp>

// ---- linkable reference helpers (external linkage in the reference TUs) -------------------------------------------------
float area(glm::vec3 v0, glm::vec3 v1, glm::vec3 v2);                                              // ray_tracing.cpp:17-21
void sortTrianglesByCentres(std::vector<Triangle>& triangles, Mesh& onlyMesh, int longestAxis);    // bvh.cpp:122-134
void sortMeshesByCentres(std::vector<Mesh>& meshes, int longestAxis);                              // bvh.cpp:88-110
AxisAlignedBox getBoundingBoxFromMeshes(std::vector<Mesh>& meshes);                                // bvh.cpp:235-268
bool intersectDataStructure(Ray& ray, HitInfo& hitInfo, const Node& root, const std::vector<Node>& nodes); // bvh.cpp:831-844

// draw.h hooks referenced by the BVH TU (debug drawing is GL-only, out of scope): stubs.
void drawAABB(const AxisAlignedBox&, DrawMode, const glm::vec3&, float) {}

// ---- --wrap interposers: count tests and remember the last accepted triangle ---------------------------------------------
struct WrapState {
    uint64_t nBox = 0, nTri = 0;
    const glm::vec3* v[3] = {nullptr, nullptr, nullptr};
    bool haveTri = false;
};
static thread_local WrapState g_ws;

extern "C" {
bool __real__Z24intersectRayWithTriangleRKN3glm4vec3ES2_S2_R3RayR7HitInfoS2_S2_S2_(
    const glm::vec3&, const glm::vec3&, const glm::vec3&, Ray&, HitInfo&, const glm::vec3&, const glm::vec3&, const glm::vec3&);
bool __real__Z21intersectRayWithShapeRK14AxisAlignedBoxR3Ray(const AxisAlignedBox&, Ray&);
}
extern "C" bool __wrap__Z24intersectRayWithTriangleRKN3glm4vec3ES2_S2_R3RayR7HitInfoS2_S2_S2_(
    const glm::vec3& v0, const glm::vec3& v1, const glm::vec3& v2, Ray& ray, HitInfo& hi, const glm::vec3& n1,
    const glm::vec3& n2, const glm::vec3& n3)
{
    g_ws.nTri++;
    bool r = __real__Z24intersectRayWithTriangleRKN3glm4vec3ES2_S2_R3RayR7HitInfoS2_S2_S2_(v0, v1, v2, ray, hi, n1, n2, n3);
    if (r) {
        g_ws.v[0] = &v0; g_ws.v[1] = &v1; g_ws.v[2] = &v2;
        g_ws.haveTri = true;
    }
    return r;
}
extern "C" bool __wrap__Z21intersectRayWithShapeRK14AxisAlignedBoxR3Ray(const AxisAlignedBox& box, Ray& ray)
{
    g_ws.nBox++;
    return __real__Z21intersectRayWithShapeRK14AxisAlignedBoxR3Ray(box, ray);
}

// ---- flat scene description shared with the product C-ABI (same field order as include/cgrt_b200.h cgrt_scene_desc) -------
struct SceneDesc {
    int32_t n_meshes;
    const int32_t* mesh_vertex_count;
    const int32_t* mesh_triangle_count;
    const float* vertices;      // [sum v][6] p.xyz n.xyz          (mesh.h:12-15)
    const uint32_t* triangles;  // [sum t][3] mesh-local indices  (mesh.h:25)
    const float* materials;     // [n_meshes][8] kd ks shininess transparency (mesh.h:17-23)
    int32_t n_spheres;
    const float* spheres;       // [n_spheres][12] center radius material(8) (scene.h:36-40)
};

struct CameraDesc { // trackball.h:47-53 private state + window aspect (window.cpp:334-337)
    float fovy, aspect, dist;
    float lookAt[3];
    float euler[3];
};

typedef std::array<uint32_t, 9> PosKey;
static PosKey keyOf(const glm::vec3& a, const glm::vec3& b, const glm::vec3& c)
{
    PosKey k;
    const glm::vec3* p[3] = {&a, &b, &c};
    for (int i = 0; i < 3; i++) {
        std::memcpy(&k[3 * i + 0], &p[i]->x, 4);
        std::memcpy(&k[3 * i + 1], &p[i]->y, 4);
        std::memcpy(&k[3 * i + 2], &p[i]->z, 4);
    }
    return k;
}

struct RefScene {
    Scene scene;
    std::vector<int32_t> meshTriOffset;         // global triangle id = offset[mesh] + local index
    std::map<PosKey, int32_t> triByPos;         // first (smallest) global id with these vertex positions
    std::map<std::array<uint32_t, 8>, int32_t> matByBits; // first mesh id with these material bits
    std::vector<Mesh> heldMeshes;               // used by mode 1 to restore scene.meshes
};

struct RefBVH {
    RefScene* rs;
    BoundingVolumeHierarchy* bvh;
};

static int32_t lookupTri(const RefScene* rs)
{
    if (!g_ws.haveTri) return -1;
    auto it = rs->triByPos.find(keyOf(*g_ws.v[0], *g_ws.v[1], *g_ws.v[2]));
    return it == rs->triByPos.end() ? -2 : it->second;
}

// ---- range-based builder that fills reference Node structs (mode 1) ------------------------------------------------------
// Follows bounding_volume_hierarchy.cpp:280-331 (getSubNodes) / :343-372 (createTree, BFS numbering) but inner nodes do
// not carry mesh copies (the verbatim traversal never reads them: intersectRecursive :748-758 only touches AABB/indices
// of inner nodes and `meshes` of leaves).
struct BuildItem {
    // either a list of whole meshes (ids, in current order) or one mesh fragment (mesh id + its current triangle list)
    std::vector<int> meshIds;
    std::vector<Triangle> tris; // only when meshIds.size()==1
};

static AxisAlignedBox boxOf(const std::vector<Mesh>& all, const BuildItem& it)
{
    // getBoundingBoxFromMeshes bvh.cpp:235-268 : seeded from first triangle's first vertex of the first mesh
    const Mesh& m0 = all[it.meshIds[0]];
    const std::vector<Triangle>& t0 = it.meshIds.size() == 1 ? it.tris : m0.triangles;
    float firstTriangleVertex = t0[0].x;
    float min_x, max_x, min_y, max_y, min_z, max_z;
    min_x = max_x = m0.vertices[firstTriangleVertex].p.x;
    min_y = max_y = m0.vertices[firstTriangleVertex].p.y;
    min_z = max_z = m0.vertices[firstTriangleVertex].p.z;
    for (size_t mi = 0; mi < it.meshIds.size(); mi++) {
        const Mesh& mesh = all[it.meshIds[mi]];
        const std::vector<Triangle>& ts = it.meshIds.size() == 1 ? it.tris : mesh.triangles;
        for (const Triangle& t : ts) {
            for (int i = 0; i < 3; i++) {
                const glm::vec3& p = mesh.vertices[(i == 0) ? t.x : ((i == 1) ? t.y : t.z)].p;
                min_x = (p.x < min_x) ? p.x : min_x;
                min_y = (p.y < min_y) ? p.y : min_y;
                min_z = (p.z < min_z) ? p.z : min_z;
                max_x = (p.x > max_x) ? p.x : max_x;
                max_y = (p.y > max_y) ? p.y : max_y;
                max_z = (p.z > max_z) ? p.z : max_z;
            }
        }
    }
    return AxisAlignedBox{glm::vec3{min_x, min_y, min_z}, glm::vec3{max_x, max_y, max_z}};
}

static bool itemIsSingleTri(const std::vector<Mesh>& all, const BuildItem& it)
{
    if (it.meshIds.size() != 1) return false;
    return it.tris.size() == 1;
}

static void fastBuild(RefScene* rs, std::vector<Node>& nodes, int maxDepth)
{
    std::vector<Mesh>& all = rs->scene.meshes;
    std::vector<BuildItem> items; // parallel to nodes
    BuildItem root;
    for (int i = 0; i < (int)all.size(); i++) root.meshIds.push_back(i);
    if (all.size() == 1) root.tris = all[0].triangles;
    bool rootLeaf = maxDepth - 1 == 0 || itemIsSingleTri(all, root); // bvh.cpp:58
    nodes.push_back(Node{rootLeaf, 0, boxOf(all, root), {}, {}});
    items.push_back(std::move(root));
    for (size_t cur = 0; cur < nodes.size(); cur++) { // createTree bvh.cpp:343-372 (BFS, children adjacent)
        if (nodes[cur].isLeaf) continue;
        const AxisAlignedBox bb = nodes[cur].AABB;
        float x = bb.upper.x - bb.lower.x, y = bb.upper.y - bb.lower.y, z = bb.upper.z - bb.lower.z;
        int longestAxis = (x > y) ? ((x > z) ? 0 : 2) : ((y > z) ? 1 : 2); // bvh.cpp:286-289
        BuildItem L, R;
        BuildItem& it = items[cur];
        if (it.meshIds.size() > 1) {
            // getChildMeshesMultipleMeshes bvh.cpp:168-179 using the reference's own sortMeshesByCentres on copies
            std::vector<Mesh> copy;
            for (int id : it.meshIds) copy.push_back(all[id]);
            // tag each copy with its id through the (otherwise unused here) transparency slot? No: keep results exact by
            // sorting an index array with the same comparator outcome instead -> run the reference sort on the copies and
            // recover ids by matching the vertex storage size + first vertex bits + triangle count.
            sortMeshesByCentres(copy, longestAxis);
            std::vector<int> sortedIds;
            std::vector<bool> used(it.meshIds.size(), false);
            for (const Mesh& m : copy) {
                for (size_t k = 0; k < it.meshIds.size(); k++) {
                    if (used[k]) continue;
                    const Mesh& o = all[it.meshIds[k]];
                    if (o.vertices.size() == m.vertices.size() && o.triangles.size() == m.triangles.size() &&
                        std::memcmp(o.vertices.data(), m.vertices.data(), o.vertices.size() * sizeof(Vertex)) == 0 &&
                        std::memcmp(&o.material, &m.material, sizeof(Material)) == 0) {
                        used[k] = true;
                        sortedIds.push_back(it.meshIds[k]);
                        break;
                    }
                }
            }
            size_t half = sortedIds.size() / 2;
            L.meshIds.assign(sortedIds.begin(), sortedIds.begin() + half);
            R.meshIds.assign(sortedIds.begin() + half, sortedIds.end());
            if (L.meshIds.size() == 1) L.tris = all[L.meshIds[0]].triangles;
            if (R.meshIds.size() == 1) R.tris = all[R.meshIds[0]].triangles;
        } else {
            // getChildMeshesOneMesh bvh.cpp:192-207 with the reference's own sortTrianglesByCentres
            std::vector<Triangle> triangles = it.tris;
            sortTrianglesByCentres(triangles, all[it.meshIds[0]], longestAxis);
            size_t half = triangles.size() / 2;
            L.meshIds = it.meshIds;
            R.meshIds = it.meshIds;
            L.tris.assign(triangles.begin(), triangles.begin() + half);
            R.tris.assign(triangles.begin() + half, triangles.end());
        }
        bool areLeaf = (nodes[cur].level + 1 == maxDepth - 1); // bvh.cpp:320
        bool lLeaf = areLeaf || itemIsSingleTri(all, L);
        bool rLeaf = areLeaf || itemIsSingleTri(all, R);
        int level = nodes[cur].level + 1;
        int lastIndex = (int)nodes.size();
        nodes[cur].indices.push_back(lastIndex);
        nodes[cur].indices.push_back(lastIndex + 1);
        nodes.push_back(Node{lLeaf, level, boxOf(all, L), {}, {}});
        nodes.push_back(Node{rLeaf, level, boxOf(all, R), {}, {}});
        items[cur] = BuildItem(); // free
        items.push_back(std::move(L));
        items.push_back(std::move(R));
    }
    // materialise compact meshes for the leaves only (vertices re-indexed; same triangle order, same Vertex values)
    for (size_t i = 0; i < nodes.size(); i++) {
        if (!nodes[i].isLeaf) continue;
        const BuildItem& it = items[i];
        for (size_t mi = 0; mi < it.meshIds.size(); mi++) {
            const Mesh& src = all[it.meshIds[mi]];
            const std::vector<Triangle>& ts = it.meshIds.size() == 1 ? it.tris : src.triangles;
            Mesh m;
            m.material = src.material;
            m.vertices.reserve(ts.size() * 3);
            m.triangles.reserve(ts.size());
            for (const Triangle& t : ts) {
                unsigned b = (unsigned)m.vertices.size();
                m.vertices.push_back(src.vertices[t.x]);
                m.vertices.push_back(src.vertices[t.y]);
                m.vertices.push_back(src.vertices[t.z]);
                m.triangles.push_back(Triangle(b, b + 1, b + 2));
            }
            nodes[i].meshes.push_back(std::move(m));
        }
    }
}

// ---- restated camera (trackball.cpp:70-73, 92-103; glm/gtc/quaternion 0.9.9.8 from memory -> "parity unpinned") -----------
struct Quat { float w, x, y, z; };
static Quat quatFromEuler(const glm::vec3& e)
{
    glm::vec3 h = e * 0.5f;
    glm::vec3 c(std::cos(h.x), std::cos(h.y), std::cos(h.z));
    glm::vec3 s(std::sin(h.x), std::sin(h.y), std::sin(h.z));
    Quat q;
    q.w = c.x * c.y * c.z + s.x * s.y * s.z;
    q.x = s.x * c.y * c.z - c.x * s.y * s.z;
    q.y = c.x * s.y * c.z + s.x * c.y * s.z;
    q.z = c.x * c.y * s.z - s.x * s.y * c.z;
    return q;
}
static glm::vec3 rotate(const Quat& q, const glm::vec3& v)
{
    const glm::vec3 QuatVector(q.x, q.y, q.z);
    const glm::vec3 uv(glm::cross(QuatVector, v));
    const glm::vec3 uuv(glm::cross(QuatVector, uv));
    return v + ((uv * q.w) + uuv) * 2.0f;
}
struct Camera {
    float fovy, aspect, dist;
    glm::vec3 lookAt, euler;
    glm::vec3 position() const { return lookAt + rotate(quatFromEuler(euler), glm::vec3(0, 0, -dist)); } // trackball.cpp:70-73
    Ray generateRay(float px, float py) const // trackball.cpp:92-103
    {
        const float halfScreenPlaceHeight = std::tan(fovy / 2.0f);
        const float halfScreenPlaceWidth = aspect * halfScreenPlaceHeight;
        const glm::vec3 cameraSpaceDirection =
            glm::normalize(glm::vec3(-px * halfScreenPlaceWidth, py * halfScreenPlaceHeight, 1.0f));
        Ray ray;
        ray.origin = position();
        ray.direction = rotate(quatFromEuler(euler), cameraSpaceDirection);
        ray.t = std::numeric_limits<float>::max();
        return ray;
    }
};

// ---- restated shading recursion (main.cpp:61-310), point lights only ----------------------------------------------------
struct RenderCounters { uint64_t primary = 0, primaryHit = 0, shadow = 0, bounce = 0, nBox = 0, nTri = 0; };
struct Tracer {
    const Scene* scene;
    const BoundingVolumeHierarchy* bvh;
    int traceLimit;
    bool duplicateShading; // main.cpp:284 evaluates shading() once more and discards it
    RenderCounters* rc;
    bool countShadow;

    glm::vec3 specularOneLight(Ray& ray, const PointLight& light, const glm::vec3& fromPosToLight, HitInfo& hitInfo) const
    { // main.cpp:61-82
        glm::vec3 fromCamToPos = ray.direction;
        glm::vec3 reflected = glm::normalize(glm::reflect(fromCamToPos, hitInfo.normal));
        float specularCos = glm::dot(reflected, fromPosToLight);
        if (specularCos <= 0) return glm::vec3(0);
        glm::vec3 result = light.color * hitInfo.material.ks * std::pow(specularCos, hitInfo.material.shininess);
        return result;
    }
    glm::vec3 diffuseOneLight(Ray&, const PointLight& light, const glm::vec3& fromPosToLight, HitInfo& hitInfo) const
    { // main.cpp:84-98
        float diffuseCos = glm::dot(fromPosToLight, hitInfo.normal);
        if (diffuseCos <= 0) return glm::vec3(0);
        return light.color * hitInfo.material.kd * diffuseCos;
    }
    bool pointInShadow(glm::vec3& pointOn, const PointLight& light) const
    { // main.cpp:104-135
        glm::vec3 fromPosToLight = light.position - pointOn;
        Ray ray{pointOn, glm::normalize(fromPosToLight), std::numeric_limits<float>::max()};
        float epsilon = 0.001;
        ray.origin += epsilon * ray.direction;
        HitInfo shadowRayHitInfo;
        if (countShadow) rc->shadow++;
        if (bvh->intersect(ray, shadowRayHitInfo)) {
            if (ray.t + epsilon >= glm::length(fromPosToLight)) return false;
            return true;
        }
        return false;
    }
    glm::vec3 shading(Ray& ray, HitInfo& hitInfo) const
    { // main.cpp:160-235, point-light loop :220-232 (spherical lights out of scope)
        glm::vec3 pointOn = ray.origin + ray.direction * ray.t;
        glm::vec3 result(0.0f);
        for (const PointLight& light : scene->pointLights) {
            const glm::vec3 fromPosToLight = glm::normalize(light.position - pointOn);
            if (pointInShadow(pointOn, light)) continue;
            glm::vec3 diffuse = diffuseOneLight(ray, light, fromPosToLight, hitInfo);
            glm::vec3 specular = specularOneLight(ray, light, fromPosToLight, hitInfo);
            result += diffuse;
            result += specular;
        }
        return result;
    }
    void shade(int level, Ray ray, glm::vec3& color, HitInfo& hitInfo)
    { // main.cpp:241-264
        countShadow = true;
        glm::vec3 directColor = shading(ray, hitInfo);
        if (hitInfo.material.ks.z <= 0.01f) { // comma operator: only the last operand decides (main.cpp:246)
            color = directColor;
            return;
        }
        glm::vec3 fromCamToPos = ray.direction;
        glm::vec3 reflected = glm::normalize(glm::reflect(fromCamToPos, hitInfo.normal));
        Ray reflectedRay = {ray.origin + ray.direction * ray.t, reflected, glm::length(fromCamToPos)};
        float epsilon = 0.001;
        reflectedRay.origin += epsilon * reflectedRay.direction;
        glm::vec3 reflectedColor;
        trace(level + 1, reflectedRay, reflectedColor);
        color = directColor + reflectedColor * hitInfo.material.ks;
    }
    void trace(int level, Ray ray, glm::vec3& color)
    { // main.cpp:265-295
        if (level >= traceLimit) {
            color = glm::vec3(0.0f);
            return;
        }
        if (level == 0) rc->primary++; else rc->bounce++;
        HitInfo hitInfo;
        if (bvh->intersect(ray, hitInfo)) {
            if (level == 0) rc->primaryHit++;
            if (duplicateShading) {
                countShadow = false; // rays are counted once, the work is still done
                glm::vec3 shadingResult = shading(ray, hitInfo);
                (void)shadingResult;
            }
            shade(level, ray, color, hitInfo);
        } else {
            color = glm::vec3(0.0f);
        }
    }
};

extern "C" {

void* ref_scene_create(const SceneDesc* d)
{
    RefScene* rs = new RefScene();
    size_t vo = 0, to = 0;
    int32_t gid = 0;
    for (int m = 0; m < d->n_meshes; m++) {
        Mesh mesh;
        int nv = d->mesh_vertex_count[m], nt = d->mesh_triangle_count[m];
        mesh.vertices.resize(nv);
        for (int i = 0; i < nv; i++) {
            const float* v = d->vertices + 6 * (vo + i);
            mesh.vertices[i] = Vertex{glm::vec3(v[0], v[1], v[2]), glm::vec3(v[3], v[4], v[5])};
        }
        for (int i = 0; i < nt; i++) {
            const uint32_t* t = d->triangles + 3 * (to + i);
            mesh.triangles.push_back(Triangle(t[0], t[1], t[2]));
        }
        const float* mt = d->materials + 8 * m;
        mesh.material.kd = glm::vec3(mt[0], mt[1], mt[2]);
        mesh.material.ks = glm::vec3(mt[3], mt[4], mt[5]);
        mesh.material.shininess = mt[6];
        mesh.material.transparency = mt[7];
        rs->meshTriOffset.push_back(gid);
        for (int i = 0; i < nt; i++) {
            const Triangle& t = mesh.triangles[i];
            PosKey k = keyOf(mesh.vertices[t.x].p, mesh.vertices[t.y].p, mesh.vertices[t.z].p);
            rs->triByPos.emplace(k, gid + i); // keeps the first (smallest) id
        }
        std::array<uint32_t, 8> mk;
        std::memcpy(mk.data(), mt, 32);
        rs->matByBits.emplace(mk, m);
        gid += nt;
        vo += nv;
        to += nt;
        rs->scene.meshes.push_back(std::move(mesh));
    }
    for (int s = 0; s < d->n_spheres; s++) {
        const float* p = d->spheres + 12 * s;
        Sphere sp;
        sp.center = glm::vec3(p[0], p[1], p[2]);
        sp.radius = p[3];
        sp.material.kd = glm::vec3(p[4], p[5], p[6]);
        sp.material.ks = glm::vec3(p[7], p[8], p[9]);
        sp.material.shininess = p[10];
        sp.material.transparency = p[11];
        rs->scene.spheres.push_back(sp);
    }
    return rs;
}
void ref_scene_destroy(void* h) { delete (RefScene*)h; }

void ref_scene_set_lights(void* h, int n, const float* posrgb)
{
    RefScene* rs = (RefScene*)h;
    rs->scene.pointLights.clear();
    for (int i = 0; i < n; i++)
        rs->scene.pointLights.push_back(PointLight{glm::vec3(posrgb[6 * i], posrgb[6 * i + 1], posrgb[6 * i + 2]),
                                                   glm::vec3(posrgb[6 * i + 3], posrgb[6 * i + 4], posrgb[6 * i + 5])});
}

// mode 0: the reference constructor, verbatim. mode 1: range-based fill of reference Node structs (maxDepth as given).
void* ref_bvh_create(void* h, int mode, int maxDepth)
{
    RefScene* rs = (RefScene*)h;
    RefBVH* b = new RefBVH();
    b->rs = rs;
    if (mode == 0) {
        b->bvh = new BoundingVolumeHierarchy(&rs->scene); // maxDepth is the reference literal 12 (bvh.cpp:48)
    } else {
        std::vector<Mesh> keep;
        keep.swap(rs->scene.meshes);
        b->bvh = new BoundingVolumeHierarchy(&rs->scene); // empty scene: constructor returns early (bvh.cpp:52-55)
        keep.swap(rs->scene.meshes);
        b->bvh->maxDepth = maxDepth;
        if (!rs->scene.meshes.empty()) fastBuild(rs, b->bvh->nodes, maxDepth);
    }
    return b;
}
void ref_bvh_destroy(void* h)
{
    RefBVH* b = (RefBVH*)h;
    delete b->bvh;
    delete b;
}
int ref_bvh_num_levels(void* h) { return ((RefBVH*)h)->bvh->numLevels(); }
int ref_bvh_num_nodes(void* h) { return (int)((RefBVH*)h)->bvh->nodes.size(); }

// meta[n][5] = isLeaf, level, child0, child1, nTriangles(leaf only) ; aabb[n][6] = lower, upper
void ref_bvh_export_nodes(void* h, int32_t* meta, float* aabb)
{
    RefBVH* b = (RefBVH*)h;
    const std::vector<Node>& nodes = b->bvh->nodes;
    for (size_t i = 0; i < nodes.size(); i++) {
        const Node& n = nodes[i];
        int nt = 0;
        if (n.isLeaf)
            for (const Mesh& m : n.meshes) nt += (int)m.triangles.size();
        meta[5 * i + 0] = n.isLeaf ? 1 : 0;
        meta[5 * i + 1] = n.level;
        meta[5 * i + 2] = n.indices.size() > 0 ? n.indices[0] : -1;
        meta[5 * i + 3] = n.indices.size() > 1 ? n.indices[1] : -1;
        meta[5 * i + 4] = nt;
        aabb[6 * i + 0] = n.AABB.lower.x; aabb[6 * i + 1] = n.AABB.lower.y; aabb[6 * i + 2] = n.AABB.lower.z;
        aabb[6 * i + 3] = n.AABB.upper.x; aabb[6 * i + 4] = n.AABB.upper.y; aabb[6 * i + 5] = n.AABB.upper.z;
    }
}
// canonical global triangle ids of a leaf in the leaf's own visiting order (intersectLeaf bvh.cpp:535-553)
int ref_bvh_leaf_triangles(void* h, int node, int32_t* out, int cap)
{
    RefBVH* b = (RefBVH*)h;
    const Node& n = b->bvh->nodes[node];
    int k = 0;
    for (const Mesh& m : n.meshes)
        for (const Triangle& t : m.triangles) {
            auto it = b->rs->triByPos.find(keyOf(m.vertices[t.x].p, m.vertices[t.y].p, m.vertices[t.z].p));
            if (k < cap) out[k] = it == b->rs->triByPos.end() ? -2 : it->second;
            k++;
        }
    return k;
}

static void fillHit(const RefScene* rs, const Ray& in, const Ray& ray, bool hit, const HitInfo& hi, float* out)
{
    // out[8] = t, triId(bits), alpha, beta, gamma, normal.xyz
    int32_t id = -1;
    float a = 0, be = 0, g = 0;
    glm::vec3 n(0.0f);
    if (hit && g_ws.haveTri) {
        id = lookupTri(rs);
        // barycentrics exactly as ray_tracing.cpp:94-96 evaluates them for the accepted hit
        const glm::vec3 &v0 = *g_ws.v[0], &v1 = *g_ws.v[1], &v2 = *g_ws.v[2];
        glm::vec3 p = ray.origin + ray.direction * ray.t;
        a = area(p, v1, v2) / area(v0, v1, v2);
        be = area(p, v0, v2) / area(v0, v1, v2);
        g = area(p, v0, v1) / area(v0, v1, v2);
    }
    if (hit) n = hi.normal;
    out[0] = hit ? ray.t : in.t;
    std::memcpy(&out[1], &id, 4);
    out[2] = a; out[3] = be; out[4] = g;
    out[5] = n.x; out[6] = n.y; out[7] = n.z;
}

// rays[n][8] = origin.xyz, t, direction.xyz, pad.  hits[n][8] see fillHit.  counts[n][2] = box tests, triangle tests (or null)
// NOTE sphere hits leave triId = id of the last accepted triangle or -1 (bvh.cpp:878-879 does not touch the material either).
void ref_intersect(void* h, const float* rays, int64_t n, float* hits, uint32_t* counts, int nthreads)
{
    RefBVH* b = (RefBVH*)h;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 256)
#endif
    for (int64_t i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        Ray in{glm::vec3(r[0], r[1], r[2]), glm::vec3(r[4], r[5], r[6]), r[3]};
        Ray ray = in;
        HitInfo hi;
        g_ws = WrapState();
        bool hit = b->bvh->intersect(ray, hi);
        fillHit(b->rs, in, ray, hit, hi, hits + 8 * i);
        if (counts) {
            counts[2 * i] = (uint32_t)g_ws.nBox;
            counts[2 * i + 1] = (uint32_t)g_ws.nTri;
        }
    }
}

// brute force over all meshes in scene order: intersectRayWithShape(const Mesh&, ...) ray_tracing.cpp:202-213
void ref_intersect_brute(void* h, const float* rays, int64_t n, float* hits, int nthreads)
{
    RefScene* rs = (RefScene*)h;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
#endif
    for (int64_t i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        Ray in{glm::vec3(r[0], r[1], r[2]), glm::vec3(r[4], r[5], r[6]), r[3]};
        Ray ray = in;
        HitInfo hi;
        g_ws = WrapState();
        bool hit = false;
        for (const Mesh& m : rs->scene.meshes) {
            // same loop as ray_tracing.cpp:202-213 but routed through the wrapped symbol so ids are recorded
            for (const auto& tri : m.triangles) {
                const auto& v0 = m.vertices[tri[0]];
                const auto& v1 = m.vertices[tri[1]];
                const auto& v2 = m.vertices[tri[2]];
                hit |= __wrap__Z24intersectRayWithTriangleRKN3glm4vec3ES2_S2_R3RayR7HitInfoS2_S2_S2_(
                    v0.p, v1.p, v2.p, ray, hi, v0.n, v1.n, v2.n);
            }
        }
        fillHit(rs, in, ray, hit, hi, hits + 8 * i);
    }
}

// the reference's own brute-force entry, un-instrumented (returns t only) — used to cross-check the instrumented loop above
void ref_intersect_mesh_verbatim(void* h, const float* rays, int64_t n, float* tOut, uint8_t* hitOut)
{
    RefScene* rs = (RefScene*)h;
    for (int64_t i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        Ray ray{glm::vec3(r[0], r[1], r[2]), glm::vec3(r[4], r[5], r[6]), r[3]};
        HitInfo hi;
        bool hit = false;
        for (const Mesh& m : rs->scene.meshes) hit |= intersectRayWithShape(m, ray, hi);
        tOut[i] = ray.t;
        hitOut[i] = hit;
    }
}

// ---- unit entry points: one call per element of a batch, straight into the verbatim functions ----------------------------
// boxes[n][6], rays[n][8] -> hit[n], t[n]  (ray_tracing.cpp:162-200)
void ref_ray_aabb(const float* boxes, const float* rays, int64_t n, uint8_t* hit, float* t)
{
    for (int64_t i = 0; i < n; i++) {
        const float* b = boxes + 6 * i;
        const float* r = rays + 8 * i;
        AxisAlignedBox box{glm::vec3(b[0], b[1], b[2]), glm::vec3(b[3], b[4], b[5])};
        Ray ray{glm::vec3(r[0], r[1], r[2]), glm::vec3(r[4], r[5], r[6]), r[3]};
        hit[i] = __real__Z21intersectRayWithShapeRK14AxisAlignedBoxR3Ray(box, ray);
        t[i] = ray.t;
    }
}
// tris[n][18] = v0 v1 v2 n0 n1 n2 ; out[n][8] = t, hit(int32), alpha,beta,gamma, normal  (ray_tracing.cpp:86-114)
void ref_ray_triangle(const float* tris, const float* rays, int64_t n, float* out)
{
    for (int64_t i = 0; i < n; i++) {
        const float* q = tris + 18 * i;
        const float* r = rays + 8 * i;
        glm::vec3 v0(q[0], q[1], q[2]), v1(q[3], q[4], q[5]), v2(q[6], q[7], q[8]);
        glm::vec3 n0(q[9], q[10], q[11]), n1(q[12], q[13], q[14]), n2(q[15], q[16], q[17]);
        Ray ray{glm::vec3(r[0], r[1], r[2]), glm::vec3(r[4], r[5], r[6]), r[3]};
        HitInfo hi;
        hi.normal = glm::vec3(0.0f);
        bool hit = __real__Z24intersectRayWithTriangleRKN3glm4vec3ES2_S2_R3RayR7HitInfoS2_S2_S2_(v0, v1, v2, ray, hi, n0, n1, n2);
        float* o = out + 8 * i;
        int32_t h32 = hit ? 1 : 0;
        o[0] = ray.t;
        std::memcpy(&o[1], &h32, 4);
        o[2] = o[3] = o[4] = 0.0f;
        if (hit) {
            glm::vec3 p = ray.origin + ray.direction * ray.t;
            o[2] = area(p, v1, v2) / area(v0, v1, v2);
            o[3] = area(p, v0, v2) / area(v0, v1, v2);
            o[4] = area(p, v0, v1) / area(v0, v1, v2);
        }
        o[5] = hi.normal.x; o[6] = hi.normal.y; o[7] = hi.normal.z;
    }
}
// planes[n][4] = normal.xyz, D ; (ray_tracing.cpp:40-72)
void ref_ray_plane(const float* planes, const float* rays, int64_t n, uint8_t* hit, float* t)
{
    for (int64_t i = 0; i < n; i++) {
        const float* p = planes + 4 * i;
        const float* r = rays + 8 * i;
        Plane pl;
        pl.normal = glm::vec3(p[0], p[1], p[2]);
        pl.D = p[3];
        Ray ray{glm::vec3(r[0], r[1], r[2]), glm::vec3(r[4], r[5], r[6]), r[3]};
        hit[i] = intersectRayWithPlane(pl, ray);
        t[i] = ray.t;
    }
}
// tri[n][9] -> planes[n][4]  (ray_tracing.cpp:74-82)
void ref_triangle_plane(const float* tris, int64_t n, float* planes)
{
    for (int64_t i = 0; i < n; i++) {
        const float* q = tris + 9 * i;
        Plane pl = trianglePlane(glm::vec3(q[0], q[1], q[2]), glm::vec3(q[3], q[4], q[5]), glm::vec3(q[6], q[7], q[8]));
        planes[4 * i] = pl.normal.x; planes[4 * i + 1] = pl.normal.y; planes[4 * i + 2] = pl.normal.z; planes[4 * i + 3] = pl.D;
    }
}
// in[n][15] = v0 v1 v2 n p  (ray_tracing.cpp:23-38)
void ref_point_in_triangle(const float* in, int64_t n, uint8_t* inside)
{
    for (int64_t i = 0; i < n; i++) {
        const float* q = in + 15 * i;
        inside[i] = pointInTriangle(glm::vec3(q[0], q[1], q[2]), glm::vec3(q[3], q[4], q[5]), glm::vec3(q[6], q[7], q[8]),
                                    glm::vec3(q[9], q[10], q[11]), glm::vec3(q[12], q[13], q[14]));
    }
}
// spheres[n][4] = center, radius ; out[n][5] = t, hit(int32), normal   (ray_tracing.cpp:118-158)
void ref_ray_sphere(const float* spheres, const float* rays, int64_t n, float* out)
{
    for (int64_t i = 0; i < n; i++) {
        const float* s = spheres + 4 * i;
        const float* r = rays + 8 * i;
        Sphere sp;
        sp.center = glm::vec3(s[0], s[1], s[2]);
        sp.radius = s[3];
        Ray ray{glm::vec3(r[0], r[1], r[2]), glm::vec3(r[4], r[5], r[6]), r[3]};
        HitInfo hi;
        hi.normal = glm::vec3(0.0f);
        bool hit = intersectRayWithShape(sp, ray, hi);
        int32_t h32 = hit ? 1 : 0;
        float* o = out + 5 * i;
        o[0] = ray.t;
        std::memcpy(&o[1], &h32, 4);
        o[2] = hi.normal.x; o[3] = hi.normal.y; o[4] = hi.normal.z;
    }
}

// primary rays exactly as main.cpp:691-694 + trackball.cpp:92-103 produce them; rays[W*H][8], pixel (x,y) at y*W+x
void ref_generate_rays(const CameraDesc* c, int W, int H, float* rays)
{
    Camera cam{c->fovy, c->aspect, c->dist, glm::vec3(c->lookAt[0], c->lookAt[1], c->lookAt[2]),
               glm::vec3(c->euler[0], c->euler[1], c->euler[2])};
    for (int y = 0; y < H; y++)
        for (int x = 0; x != W; x++) {
            float px = float(x) / W * 2.0f - 1.0f;
            float py = float(y) / H * 2.0f - 1.0f;
            Ray r = cam.generateRay(px, py);
            float* o = rays + 8 * ((size_t)y * W + x);
            o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z; o[3] = r.t;
            o[4] = r.direction.x; o[5] = r.direction.y; o[6] = r.direction.z; o[7] = 0.0f;
        }
}

// rgb[H][W][3] in Screen layout (row H-1-y, screen.cpp:30-36). Rows y in [y0,y1) only (bounded samples); others untouched.
// counters[6] = primary, primaryHit, shadow, bounce, boxTests, triTests (summed over threads).
// duplicateShading=1 reproduces main.cpp:284 (the reference's real cost); rays are counted once either way.
void ref_render(void* h, const CameraDesc* c, int W, int H, int traceLimit, int duplicateShading, float* rgb,
                uint64_t* counters, int y0, int y1, int nthreads)
{
    RefBVH* b = (RefBVH*)h;
    Camera cam{c->fovy, c->aspect, c->dist, glm::vec3(c->lookAt[0], c->lookAt[1], c->lookAt[2]),
               glm::vec3(c->euler[0], c->euler[1], c->euler[2])};
    RenderCounters total;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        RenderCounters rc;
        g_ws = WrapState();
        Tracer tr{&b->rs->scene, b->bvh, traceLimit, duplicateShading != 0, &rc, true};
#ifdef _OPENMP
#pragma omp for // main.cpp:653-656: plain `omp parallel for` (static schedule) over image rows
#endif
        for (int y = y0; y < y1; y++) {
            for (int x = 0; x != W; x++) {
                float px = float(x) / W * 2.0f - 1.0f; // main.cpp:691-693
                float py = float(y) / H * 2.0f - 1.0f;
                Ray cameraRay = cam.generateRay(px, py);
                glm::vec3 color;
                tr.trace(0, cameraRay, color); // getFinalColor main.cpp:298-310
                const size_t i = (size_t)(H - 1 - y) * W + x; // Screen::setPixel screen.cpp:34
                rgb[3 * i] = color.x; rgb[3 * i + 1] = color.y; rgb[3 * i + 2] = color.z;
            }
        }
        rc.nBox = g_ws.nBox;
        rc.nTri = g_ws.nTri;
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            total.primary += rc.primary; total.primaryHit += rc.primaryHit; total.shadow += rc.shadow;
            total.bounce += rc.bounce; total.nBox += rc.nBox; total.nTri += rc.nTri;
        }
    }
    if (counters) {
        counters[0] = total.primary; counters[1] = total.primaryHit; counters[2] = total.shadow;
        counters[3] = total.bounce; counters[4] = total.nBox; counters[5] = total.nTri;
    }
}

int ref_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"
