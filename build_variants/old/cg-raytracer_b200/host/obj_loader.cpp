// OBJ/MTL import + scene presets: the data source of the hot path (src/mesh.cpp:58-166, src/scene.cpp:4-69).
// The reference delegates parsing to assimp 5.0.1 (ReadFile with aiProcess_GenNormals | aiProcess_Triangulate,
// mesh.cpp:65-66), which is not vendored in the reference tree and not installable here. This file restates the parts of
// assimp's OBJ importer that decide the Mesh list the path consumes ("parity unpinned": from the published algorithm, checked
// only structurally — triangle / vertex / mesh counts and BVH levels of the bundled scenes match the report):
//   * ObjFileParser: `v`, `vn`, `f` (v, v/vt, v//vn, v/vt/vn, negative indices), `o`, `g` (groups map to objects),
//     `usemtl` (a new sub-mesh only when the current one already has faces and a different material), `mtllib`
//   * every face corner becomes its own vertex (no index sharing), one Mesh per (object, material run) with >= 1 face
//   * objects are visited in REVERSE file order: mesh.cpp walks the node tree with an explicit stack (:75-79, :131-133)
//   * TriangulateProcess: quads split (0,1,2)(0,2,3) from the first concave corner, larger polygons fanned (assimp ear-clips
//     those; none of the bundled scenes has one)
//   * GenFaceNormalsProcess (only for meshes without normals): normalize((v1-v0) x (v_last-v0)) written to all corners,
//     faces processed in order so shared quad corners keep the second triangle's normal
//   * ObjFileMtlImporter: Kd, Ks, Ns, d; defaults kd 0.6, ks 0, Ns 0, d 1
//   * fast_atoreal_move<float>: integer part + (up to 15 fractional digits as double * 10^-n) narrowed to float, then added
#include "cgrt_host.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

namespace {

// ---- assimp fast_atof -----------------------------------------------------------------------------------------------------------
const double kFastAtofTable[16] = {0.0,     0.1,      0.01,      0.001,      0.0001,      0.00001,      0.000001,      0.0000001,
                                   0.00000001, 0.000000001, 0.0000000001, 0.00000000001, 0.000000000001, 0.0000000000001,
                                   0.00000000000001, 0.000000000000001};

uint64_t strtoul10_64(const char* in, const char** out, unsigned* maxInOut)
{
    unsigned cur = 0;
    uint64_t value = 0;
    for (;;) {
        if (*in < '0' || *in > '9') break;
        const uint64_t nv = value * 10 + (uint64_t)(*in - '0');
        if (nv < value) break; // overflow: keep what we have
        value = nv;
        ++in;
        ++cur;
        if (maxInOut && *maxInOut == cur) {
            while (*in >= '0' && *in <= '9') ++in; // skip the digits beyond the relevant ones
            if (out) *out = in;
            return value;
        }
    }
    if (out) *out = in;
    if (maxInOut) *maxInOut = cur;
    return value;
}

bool fastAtof(const char* c, float& out)
{
    float f = 0;
    const bool inv = (*c == '-');
    if (inv || *c == '+') ++c;
    if (!(c[0] >= '0' && c[0] <= '9') && !((c[0] == '.' || c[0] == ',') && c[1] >= '0' && c[1] <= '9')) return false;
    if (*c != '.' && *c != ',') f = static_cast<float>(strtoul10_64(c, &c, nullptr));
    if ((*c == '.' || *c == ',') && c[1] >= '0' && c[1] <= '9') {
        ++c;
        unsigned diff = 15;
        double pl = static_cast<double>(strtoul10_64(c, &c, &diff));
        pl *= kFastAtofTable[diff];
        f += static_cast<float>(pl);
    } else if (*c == '.') {
        ++c;
    }
    if (*c == 'e' || *c == 'E') {
        ++c;
        const bool einv = (*c == '-');
        if (einv || *c == '+') ++c;
        float ex = static_cast<float>(strtoul10_64(c, &c, nullptr));
        if (einv) ex = -ex;
        f *= std::pow(10.0f, ex);
    }
    if (inv) f = -f;
    out = f;
    return true;
}

struct V3f {
    float x = 0, y = 0, z = 0;
};
V3f sub(const V3f& a, const V3f& b) { return V3f{a.x - b.x, a.y - b.y, a.z - b.z}; }
V3f crossAi(const V3f& a, const V3f& b) { return V3f{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
float lengthAi(const V3f& v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
V3f divAi(const V3f& v, float f) // aiVector3t::operator/= multiplies by the reciprocal
{
    const float invF = 1.0f / f;
    return V3f{v.x * invF, v.y * invF, v.z * invF};
}
float dotAi(const V3f& a, const V3f& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct ObjMaterial {
    std::string name;
    float kd[3] = {0.6f, 0.6f, 0.6f};
    float ks[3] = {0.0f, 0.0f, 0.0f};
    float shininess = 0.0f;
    float alpha = 1.0f;
};

struct ObjFace {
    std::vector<int> v, n; // 0-based indices into the file-global arrays
};
struct ObjMesh {
    std::vector<ObjFace> faces;
    int material = -1; // index into materials, -1 = NoMaterial (-> material 0, the default)
    bool hasNormals = false;
};
struct ObjObject {
    std::string name;
    std::vector<int> meshes;
};

std::vector<std::string> tokens(const std::string& line)
{
    std::vector<std::string> t;
    std::istringstream ss(line);
    std::string w;
    while (ss >> w) t.push_back(w);
    return t;
}
std::string restOfLine(const std::string& line, size_t skip)
{
    size_t b = line.find_first_not_of(" \t", skip);
    if (b == std::string::npos) return "";
    size_t e = line.find_last_not_of(" \t\r\n");
    return line.substr(b, e - b + 1);
}

struct ObjModel {
    std::vector<V3f> positions, normals;
    std::vector<ObjMaterial> materials; // [0] = DefaultMaterial
    std::map<std::string, int> materialIndex;
    std::vector<ObjMesh> meshes;
    std::vector<ObjObject> objects;
    int currentObject = -1, currentMesh = -1, currentMaterial = -1;
    std::string activeGroup;

    ObjModel()
    {
        ObjMaterial d;
        d.name = "DefaultMaterial";
        materials.push_back(d);
        materialIndex[d.name] = 0;
    }
    void createMesh()
    {
        meshes.push_back(ObjMesh());
        currentMesh = (int)meshes.size() - 1;
        if (currentObject >= 0) objects[currentObject].meshes.push_back(currentMesh);
    }
    void createObject(const std::string& name)
    {
        ObjObject o;
        o.name = name;
        objects.push_back(o);
        currentObject = (int)objects.size() - 1;
        createMesh();
        if (currentMaterial >= 0) meshes[currentMesh].material = currentMaterial;
    }
    void useMaterial(const std::string& name)
    {
        if (currentMaterial >= 0 && materials[currentMaterial].name == name) return; // same material: ignored
        auto it = materialIndex.find(name);
        if (it == materialIndex.end()) { // unknown material: a new default-valued material of that name
            ObjMaterial m;
            m.name = name;
            materials.push_back(m);
            materialIndex[name] = (int)materials.size() - 1;
            currentMaterial = (int)materials.size() - 1;
        } else {
            currentMaterial = it->second;
        }
        bool needNew = false;
        if (currentMesh < 0) needNew = true;
        else {
            const ObjMesh& cm = meshes[currentMesh];
            if (cm.material != -1 && cm.material != currentMaterial && !cm.faces.empty()) needNew = true;
        }
        if (needNew) createMesh();
        meshes[currentMesh].material = currentMaterial;
    }
};

void loadMtl(const std::filesystem::path& file, ObjModel& model)
{
    std::ifstream in(file);
    if (!in) {
        std::cerr << "OBJ: unable to open material file " << file << std::endl;
        return;
    }
    std::string line;
    int cur = -1;
    auto color = [](const std::vector<std::string>& t, float out[3]) {
        float r = 0;
        if (t.size() > 1) fastAtof(t[1].c_str(), r);
        float g = r, b = r; // a single component is replicated
        if (t.size() > 3) {
            fastAtof(t[2].c_str(), g);
            fastAtof(t[3].c_str(), b);
        }
        out[0] = r; out[1] = g; out[2] = b;
    };
    while (std::getline(in, line)) {
        const std::vector<std::string> t = tokens(line);
        if (t.empty() || t[0][0] == '#') continue;
        if (t[0] == "newmtl") {
            const std::string name = restOfLine(line, line.find("newmtl") + 6);
            auto it = model.materialIndex.find(name);
            if (it == model.materialIndex.end()) {
                ObjMaterial m;
                m.name = name;
                model.materials.push_back(m);
                cur = (int)model.materials.size() - 1;
                model.materialIndex[name] = cur;
            } else {
                cur = it->second;
            }
        } else if (cur >= 0) {
            ObjMaterial& m = model.materials[cur];
            if (t[0] == "Kd") color(t, m.kd);
            else if (t[0] == "Ks") color(t, m.ks);
            else if (t[0] == "Ns" && t.size() > 1) fastAtof(t[1].c_str(), m.shininess);
            else if (t[0] == "d" && t.size() > 1) fastAtof(t[1].c_str(), m.alpha);
        }
    }
}

void parseObj(const std::filesystem::path& file, ObjModel& model)
{
    std::ifstream in(file);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        const std::vector<std::string> t = tokens(line);
        if (t.empty()) continue;
        const std::string& k = t[0];
        if (k == "v") {
            V3f p;
            if (t.size() >= 4) {
                fastAtof(t[1].c_str(), p.x);
                fastAtof(t[2].c_str(), p.y);
                fastAtof(t[3].c_str(), p.z);
                if (t.size() == 5) { // homogeneous coordinate
                    float w = 1.0f;
                    fastAtof(t[4].c_str(), w);
                    p.x /= w; p.y /= w; p.z /= w;
                }
            }
            model.positions.push_back(p);
        } else if (k == "vn") {
            V3f n;
            if (t.size() >= 4) {
                fastAtof(t[1].c_str(), n.x);
                fastAtof(t[2].c_str(), n.y);
                fastAtof(t[3].c_str(), n.z);
            }
            model.normals.push_back(n);
        } else if (k == "f") {
            ObjFace face;
            bool hasNormal = false;
            for (size_t i = 1; i < t.size(); i++) {
                // v, v/vt, v//vn, v/vt/vn ; negative = relative to the end of the arrays read so far
                int idx[3] = {0, 0, 0};
                int part = 0;
                const char* c = t[i].c_str();
                while (*c && part < 3) {
                    if (*c == '/') { part++; c++; continue; }
                    char* e = nullptr;
                    long v = std::strtol(c, &e, 10);
                    if (e == c) break;
                    idx[part] = (int)v;
                    c = e;
                }
                if (idx[0] == 0) continue;
                face.v.push_back(idx[0] > 0 ? idx[0] - 1 : (int)model.positions.size() + idx[0]);
                if (idx[2] != 0) {
                    face.n.push_back(idx[2] > 0 ? idx[2] - 1 : (int)model.normals.size() + idx[2]);
                    hasNormal = true;
                }
            }
            if (face.v.empty()) continue;
            if (model.currentObject < 0) model.createObject("defaultobject");
            if (model.currentMesh < 0) model.createMesh();
            ObjMesh& m = model.meshes[model.currentMesh];
            m.faces.push_back(face);
            if (!m.hasNormals && hasNormal) m.hasNormals = true;
        } else if (k == "o") {
            const std::string name = restOfLine(line, line.find('o') + 1);
            if (!name.empty()) {
                int found = -1;
                for (size_t i = 0; i < model.objects.size(); i++)
                    if (model.objects[i].name == name) found = (int)i;
                if (found >= 0) model.currentObject = found;
                else model.createObject(name);
            }
        } else if (k == "g") {
            const std::string name = restOfLine(line, line.find('g') + 1);
            if (model.activeGroup != name) {
                model.createObject(name); // groups are mapped into the object structure
                model.activeGroup = name;
            }
        } else if (k == "usemtl") {
            model.useMaterial(restOfLine(line, line.find("usemtl") + 6));
        } else if (k == "mtllib") {
            const std::string name = restOfLine(line, line.find("mtllib") + 6);
            loadMtl(file.parent_path() / name, model);
        }
        // `s`, `vt`, `#`, `l`, `p`: not relevant to the Mesh list the path consumes
    }
}

// TriangulateProcess + GenFaceNormalsProcess + the Mesh conversion of mesh.cpp:83-129 for one OBJ sub-mesh
bool convertMesh(const ObjModel& model, const ObjMesh& om, Mesh& out)
{
    if (om.faces.empty()) return false;
    std::vector<V3f> pos, nor;
    std::vector<std::vector<unsigned>> faces;
    const bool withNormals = !model.normals.empty() && om.hasNormals;
    for (const ObjFace& f : om.faces) {
        std::vector<unsigned> idx;
        for (size_t c = 0; c < f.v.size(); c++) {
            if (f.v[c] < 0 || f.v[c] >= (int)model.positions.size()) throw std::exception();
            idx.push_back((unsigned)pos.size());
            pos.push_back(model.positions[f.v[c]]);
            V3f n;
            if (withNormals && c < f.n.size() && f.n[c] >= 0 && f.n[c] < (int)model.normals.size()) n = model.normals[f.n[c]];
            nor.push_back(n);
        }
        faces.push_back(idx);
    }
    // ---- triangulate
    std::vector<Triangle> tris;
    for (const std::vector<unsigned>& f : faces) {
        if (f.size() < 3) continue; // points / lines: "Found a face which is not a triangle, discarding" (mesh.cpp:94-97)
        if (f.size() == 3) {
            tris.emplace_back(f[0], f[1], f[2]);
        } else if (f.size() == 4) {
            unsigned start = 0;
            for (unsigned i = 0; i < 4; ++i) { // first corner whose two adjacent angles sum above pi is concave
                const V3f& v0 = pos[f[(i + 3) % 4]];
                const V3f& v1 = pos[f[(i + 2) % 4]];
                const V3f& v2 = pos[f[(i + 1) % 4]];
                const V3f& v = pos[f[i]];
                V3f left = sub(v0, v), diag = sub(v1, v), right = sub(v2, v);
                left = divAi(left, lengthAi(left));
                diag = divAi(diag, lengthAi(diag));
                right = divAi(right, lengthAi(right));
                const float angle = std::acos(dotAi(left, diag)) + std::acos(dotAi(right, diag));
                if (angle > 3.14159265358979323846f) {
                    start = i;
                    break;
                }
            }
            tris.emplace_back(f[start], f[(start + 1) % 4], f[(start + 2) % 4]);
            tris.emplace_back(f[start], f[(start + 2) % 4], f[(start + 3) % 4]);
        } else {
            for (size_t i = 1; i + 1 < f.size(); i++) tris.emplace_back(f[0], f[i], f[i + 1]);
        }
    }
    if (tris.empty()) return false;
    // ---- face normals for meshes that came without normals
    if (!withNormals) {
        for (const Triangle& t : tris) {
            const V3f e1 = sub(pos[t.y], pos[t.x]), e2 = sub(pos[t.z], pos[t.x]);
            V3f n = crossAi(e1, e2);
            const float len = lengthAi(n);
            if (len > 0) n = divAi(n, len);
            nor[t.x] = nor[t.y] = nor[t.z] = n;
        }
    }
    out.triangles = tris;
    out.vertices.resize(pos.size());
    for (size_t i = 0; i < pos.size(); i++)
        out.vertices[i] = Vertex{glm::vec3(pos[i].x, pos[i].y, pos[i].z), glm::vec3(nor[i].x, nor[i].y, nor[i].z)};
    const ObjMaterial& mat = model.materials[om.material >= 0 ? om.material : 0];
    out.material.kd = glm::vec3(mat.kd[0], mat.kd[1], mat.kd[2]);       // AI_MATKEY_COLOR_DIFFUSE  mesh.cpp:124
    out.material.ks = glm::vec3(mat.ks[0], mat.ks[1], mat.ks[2]);       // AI_MATKEY_COLOR_SPECULAR :125
    out.material.shininess = mat.shininess;                             // AI_MATKEY_SHININESS      :126
    out.material.transparency = mat.alpha;                              // AI_MATKEY_OPACITY        :127
    return true;
}

} // namespace

// centerAndScaleToUnitMesh, src/mesh.cpp:143-166
void centerAndScaleToUnitMesh(std::vector<Mesh>& meshes)
{
    std::vector<glm::vec3> positions;
    for (const Mesh& m : meshes)
        for (const Vertex& v : m.vertices) positions.push_back(v.p);
    glm::vec3 sum(0.0f);
    for (const glm::vec3& p : positions) sum = sum + p; // std::accumulate, left fold
    const glm::vec3 center = sum / static_cast<float>(positions.size());
    float maxD = 0.0f;
    for (const glm::vec3& p : positions) {
        const glm::vec3 d = p - center;
        const float len = std::sqrt((d.x * d.x + d.y * d.y) + d.z * d.z); // glm::length
        maxD = (len < maxD) ? maxD : len;                                 // std::max(len, maxD) = (a < b) ? b : a
    }
    for (Mesh& m : meshes)
        for (Vertex& v : m.vertices) v.p = (v.p - center) / maxD;
}

std::vector<Mesh> loadMesh(const std::filesystem::path& file, bool normalize)
{
    if (!std::filesystem::exists(file)) {
        std::cerr << "File " << file << " does not exist." << std::endl;
        throw std::exception();
    }
    ObjModel model;
    parseObj(file, model);
    std::vector<Mesh> out;
    // mesh.cpp:75-79,131-133: node stack -> children (objects) are visited last-to-first; sub-meshes of a node in order
    for (int o = (int)model.objects.size() - 1; o >= 0; o--) {
        for (int mi : model.objects[o].meshes) {
            Mesh m;
            if (convertMesh(model, model.meshes[mi], m)) out.emplace_back(std::move(m));
        }
    }
    if (out.empty()) {
        std::cerr << "Assimp failed to load mesh file " << file << std::endl;
        throw std::exception();
    }
    if (normalize) centerAndScaleToUnitMesh(out);
    return out;
}

std::vector<Mesh> makeDragonStandIn(int segU, int segV)
{
    // (2,3) torus knot centre line, tube radius modulated along the knot; positions evaluated in double, stored as float.
    const double PI = 3.14159265358979323846;
    auto centre = [&](double u, double c[3]) {
        const double r = 2.0 + std::cos(3.0 * u);
        c[0] = r * std::cos(2.0 * u);
        c[1] = r * std::sin(2.0 * u);
        c[2] = -std::sin(3.0 * u);
    };
    auto surface = [&](double u, double v, double p[3]) {
        double c[3], c1[3], c0[3];
        const double h = 1e-4;
        centre(u, c);
        centre(u + h, c1);
        centre(u - h, c0);
        double t[3] = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2]};
        const double tl = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        for (double& k : t) k /= tl;
        // frame: n = normalize(c'' projected) approximated by (c1 + c0 - 2c), b = t x n
        double n[3] = {c1[0] + c0[0] - 2 * c[0], c1[1] + c0[1] - 2 * c[1], c1[2] + c0[2] - 2 * c[2]};
        const double d = n[0] * t[0] + n[1] * t[1] + n[2] * t[2];
        for (int k = 0; k < 3; k++) n[k] -= d * t[k];
        const double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (double& k : n) k /= nl;
        const double b[3] = {t[1] * n[2] - t[2] * n[1], t[2] * n[0] - t[0] * n[2], t[0] * n[1] - t[1] * n[0]};
        const double rad = 0.45 + 0.15 * std::sin(5.0 * u) + 0.05 * std::sin(7.0 * v + 3.0 * u);
        for (int k = 0; k < 3; k++) p[k] = c[k] + rad * (std::cos(v) * n[k] + std::sin(v) * b[k]);
    };
    std::vector<glm::vec3> P((size_t)segU * segV), N((size_t)segU * segV);
    for (int i = 0; i < segU; i++)
        for (int j = 0; j < segV; j++) {
            const double u = 2.0 * PI * i / segU, v = 2.0 * PI * j / segV;
            double p[3], pu[3], pv[3];
            surface(u, v, p);
            surface(u + 1e-4, v, pu);
            surface(u, v + 1e-4, pv);
            const double a[3] = {pu[0] - p[0], pu[1] - p[1], pu[2] - p[2]}, b[3] = {pv[0] - p[0], pv[1] - p[1], pv[2] - p[2]};
            double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
            const double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            P[(size_t)i * segV + j] = glm::vec3((float)p[0], (float)p[1], (float)p[2]);
            N[(size_t)i * segV + j] = glm::vec3((float)(n[0] / nl), (float)(n[1] / nl), (float)(n[2] / nl));
        }
    Mesh m;
    auto corner = [&](int i, int j) {
        const size_t k = (size_t)(i % segU) * segV + (j % segV);
        m.vertices.push_back(Vertex{P[k], N[k]});
        return (unsigned)m.vertices.size() - 1;
    };
    for (int i = 0; i < segU; i++)
        for (int j = 0; j < segV; j++) { // two triangles per quad, every corner its own vertex (as the OBJ importer does)
            unsigned a = corner(i, j), b = corner(i + 1, j), c = corner(i + 1, j + 1);
            m.triangles.emplace_back(a, b, c);
            unsigned d = corner(i, j), e = corner(i + 1, j + 1), f = corner(i, j + 1);
            m.triangles.emplace_back(d, e, f);
        }
    // harness-assigned mirror material so that the Whitted configuration has bounces to follow (assimp's default
    // material has ks = 0: a real dragon.obj without an .mtl would never reflect). Stated in every result that uses it.
    m.material.kd = glm::vec3(0.6f, 0.6f, 0.6f);
    m.material.ks = glm::vec3(0.5f, 0.5f, 0.5f);
    m.material.shininess = 32.0f;
    m.material.transparency = 1.0f;
    std::vector<Mesh> out;
    out.emplace_back(std::move(m));
    return out;
}

Scene loadScene(SceneType type, const std::filesystem::path& dataDir)
{ // src/scene.cpp:4-69
    Scene scene;
    auto add = [&](std::vector<Mesh>&& sub) {
        for (Mesh& m : sub) scene.meshes.emplace_back(std::move(m));
    };
    switch (type) {
    case SingleTriangle: {
        auto sub = loadMesh(dataDir / "triangle.obj");
        sub[0].material.kd = glm::vec3(1.0f);
        add(std::move(sub));
        scene.pointLights.push_back(PointLight{glm::vec3(-1, 1, -1), glm::vec3(1)});
    } break;
    case Cube: {
        add(loadMesh(dataDir / "cube.obj"));
        scene.pointLights.push_back(PointLight{glm::vec3(-1, 1, -1), glm::vec3(1)});
    } break;
    case CornellBox: {
        add(loadMesh(dataDir / "CornellBox-Mirror-Rotated.obj", true));
        scene.pointLights.push_back(PointLight{glm::vec3(0, 0.58f, 0), glm::vec3(1)});
    } break;
    case CornellBoxSphericalLight: {
        add(loadMesh(dataDir / "CornellBox-Mirror-Rotated.obj", true));
        scene.sphericalLight.push_back(SphericalLight{glm::vec3(0, 0.45f, 0), 0.1f, glm::vec3(1)});
    } break;
    case Monkey: {
        add(loadMesh(dataDir / "monkey-rotated.obj", true));
        scene.pointLights.push_back(PointLight{glm::vec3(-1, 1, -1), glm::vec3(1)});
        scene.pointLights.push_back(PointLight{glm::vec3(1, -1, -1), glm::vec3(1)});
    } break;
    case Dragon: {
        if (std::filesystem::exists(dataDir / "dragon.obj")) {
            add(loadMesh(dataDir / "dragon.obj", true));
        } else { // file is not part of the reference checkout: named stand-in, normalised like loadMesh(..., true) would
            auto sub = makeDragonStandIn();
            centerAndScaleToUnitMesh(sub);
            add(std::move(sub));
        }
        scene.pointLights.push_back(PointLight{glm::vec3(-1, 1, -1), glm::vec3(1)});
    } break;
    case Spheres: {
        scene.spheres.push_back(Sphere{glm::vec3(3.0f, -2.0f, 10.2f), 1.0f, Material{glm::vec3(0.8f, 0.2f, 0.2f)}});
        scene.spheres.push_back(Sphere{glm::vec3(-2.0f, 2.0f, 4.0f), 2.0f, Material{glm::vec3(0.6f, 0.8f, 0.2f)}});
        scene.spheres.push_back(Sphere{glm::vec3(0.0f, 0.0f, 6.0f), 0.75f, Material{glm::vec3(0.2f, 0.2f, 0.8f)}});
        scene.pointLights.push_back(PointLight{glm::vec3(3, 0, 3), glm::vec3(15)});
    } break;
    case Custom: {
        add(loadMesh(dataDir / "custom.obj"));
        scene.pointLights.push_back(PointLight{glm::vec3(-1, 1, -1), glm::vec3(1)});
    } break;
    }
    return scene;
}
