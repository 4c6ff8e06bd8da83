// Range-based host BVH build with the reference's split rule. See bvh_build.h.
// Parity notes (reference = src/bounding_volume_hierarchy.cpp):
//  * axis choice           :286-289  (x > y) ? ((x > z) ? 0 : 2) : ((y > z) ? 1 : 2) on the node's AABB extents
//  * single-mesh split     :192-207  std::sort of the node's triangle list by centroid[axis] with `<`, halves [0,n/2) [n/2,n)
//                                    (libstdc++ std::sort is deterministic for a given sequence + comparator outcome, so sorting
//                                    an index range with the same comparator outcome reproduces the same permutation)
//  * multi-mesh split      :168-179, :88-110  std::sort of the MESHES by the centroid[axis] of each mesh's median triangle
//  * child AABBs           :235-268  min/max via ternaries over the referenced vertices, seeded from the first triangle's
//                                    first vertex of the first mesh (index round-tripped through float, :237)
//  * leaf rule             :320-322  level+1 == maxDepth-1, or one mesh with one triangle; root :58
//  * numbering             :343-372  BFS over a growing vector, the two children pushed consecutively
#include "bvh_build.h"

#include <algorithm>
#include <functional>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstddef>
#include <limits>

namespace cgrt {
namespace {

struct Item {
    std::vector<int32_t> meshIds; // >1: whole meshes in their current order
    int32_t mesh = -1;            // ==1 mesh: fragment [begin,end) of perm[mesh]
    int32_t begin = 0, end = 0;
    bool single() const { return meshIds.empty(); }
};

struct Builder {
    const std::vector<MeshView>& meshes;
    std::vector<std::vector<float>> keys[3];     // keys[axis][mesh][tri] = centroid coordinate ((p0+p1)+p2)/3.0f
    std::vector<std::vector<int32_t>> perm;      // perm[mesh] = current triangle order
    std::vector<float> medianKey[3];             // medianKey[axis][mesh], NaN-free cache flag below
    std::vector<char> medianKnown[3];

    explicit Builder(const std::vector<MeshView>& m) : meshes(m)
    {
        const size_t nm = meshes.size();
        perm.resize(nm);
        for (int a = 0; a < 3; a++) {
            keys[a].resize(nm);
            medianKey[a].assign(nm, 0.0f);
            medianKnown[a].assign(nm, 0);
        }
        for (size_t m_ = 0; m_ < nm; m_++) {
            const MeshView& mv = meshes[m_];
            perm[m_].resize(mv.nt);
            for (int a = 0; a < 3; a++) keys[a][m_].resize(mv.nt);
            for (int32_t t = 0; t < mv.nt; t++) {
                perm[m_][t] = t;
                const uint32_t* tri = mv.triangles + 3 * (size_t)t;
                const float* p0 = mv.vertices + 6 * (size_t)tri[0];
                const float* p1 = mv.vertices + 6 * (size_t)tri[1];
                const float* p2 = mv.vertices + 6 * (size_t)tri[2];
                for (int a = 0; a < 3; a++) keys[a][m_][t] = ((p0[a] + p1[a]) + p2[a]) / 3.0f; // bvh.cpp:127-128
            }
        }
    }

    const float* pos(int32_t mesh, int32_t tri, int corner) const
    {
        const MeshView& mv = meshes[mesh];
        return mv.vertices + 6 * (size_t)mv.triangles[3 * (size_t)tri + corner];
    }

    // centroid[axis] of the median triangle of the mesh after sorting its triangles by that axis (bvh.cpp:92-99)
    float meshMedian(int32_t mesh, int axis)
    {
        if (!medianKnown[axis][mesh]) {
            std::vector<float> k = keys[axis][mesh];
            std::sort(k.begin(), k.end(), [](float a, float b) { return a < b; });
            medianKey[axis][mesh] = k[k.size() / 2];
            medianKnown[axis][mesh] = 1;
        }
        return medianKey[axis][mesh];
    }

    template <typename F>
    void forEachTriangle(const Item& it, F f) const
    {
        if (it.single()) {
            for (int32_t i = it.begin; i < it.end; i++) f(it.mesh, perm[it.mesh][i]);
        } else {
            for (int32_t m : it.meshIds)
                for (int32_t t = 0; t < meshes[m].nt; t++) f(m, perm[m][t]); // whole meshes keep their original order
        }
    }

    void boxOf(const Item& it, float lo[3], float hi[3]) const
    { // getBoundingBoxFromMeshes bvh.cpp:235-268
        const int32_t m0 = it.single() ? it.mesh : it.meshIds[0];
        const int32_t t0 = it.single() ? perm[m0][it.begin] : perm[m0][0];
        const float firstTriangleVertex = (float)meshes[m0].triangles[3 * (size_t)t0]; // bvh.cpp:237 (uint -> float -> index)
        const float* seed = meshes[m0].vertices + 6 * (size_t)firstTriangleVertex;
        float min_x, max_x, min_y, max_y, min_z, max_z;
        min_x = max_x = seed[0];
        min_y = max_y = seed[1];
        min_z = max_z = seed[2];
        forEachTriangle(it, [&](int32_t m, int32_t t) {
            for (int i = 0; i < 3; i++) {
                const float* p = pos(m, t, i);
                min_x = (p[0] < min_x) ? p[0] : min_x;
                min_y = (p[1] < min_y) ? p[1] : min_y;
                min_z = (p[2] < min_z) ? p[2] : min_z;
                max_x = (p[0] > max_x) ? p[0] : max_x;
                max_y = (p[1] > max_y) ? p[1] : max_y;
                max_z = (p[2] > max_z) ? p[2] : max_z;
            }
        });
        lo[0] = min_x; lo[1] = min_y; lo[2] = min_z;
        hi[0] = max_x; hi[1] = max_y; hi[2] = max_z;
    }

    bool singleTriangle(const Item& it) const { return it.single() && (it.end - it.begin) == 1; }

    Item wholeMesh(int32_t m) const
    {
        Item it;
        it.mesh = m;
        it.begin = 0;
        it.end = meshes[m].nt;
        return it;
    }
};

} // namespace

void buildReferenceBVH(const std::vector<MeshView>& meshes, int maxDepth, BuiltBVH& out)
{
    out.nodes.clear();
    out.leafTris.clear();
    out.numLevels = 0;
    if (meshes.empty()) return; // bvh.cpp:52-55

    Builder B(meshes);
    std::vector<Item> items; // parallel to out.nodes

    Item root;
    if (meshes.size() == 1) root = B.wholeMesh(0);
    else
        for (size_t m = 0; m < meshes.size(); m++) root.meshIds.push_back((int32_t)m);

    auto pushNode = [&](Item&& it, int level, bool isLeaf) {
        HostNode n;
        B.boxOf(it, n.lo, n.hi);
        n.child0 = n.child1 = -1;
        n.firstTri = 0;
        n.triCount = 0;
        n.level = level;
        n.isLeaf = isLeaf ? 1 : 0;
        out.nodes.push_back(n);
        items.push_back(std::move(it));
    };

    const bool rootLeaf = (maxDepth - 1 == 0) || B.singleTriangle(root); // bvh.cpp:58
    pushNode(std::move(root), 0, rootLeaf);

    for (size_t cur = 0; cur < out.nodes.size(); cur++) { // createTree bvh.cpp:343-372
        if (out.nodes[cur].isLeaf) continue;
        const HostNode nd = out.nodes[cur];
        const float x = nd.hi[0] - nd.lo[0], y = nd.hi[1] - nd.lo[1], z = nd.hi[2] - nd.lo[2];
        const int axis = (x > y) ? ((x > z) ? 0 : 2) : ((y > z) ? 1 : 2); // bvh.cpp:286-289
        Item L, R;
        Item& it = items[cur];
        if (!it.single()) {
            // getChildMeshesMultipleMeshes bvh.cpp:168-179 / sortMeshesByCentres :88-110
            std::vector<int32_t> ids = it.meshIds;
            for (int32_t m : ids) (void)B.meshMedian(m, axis);
            std::sort(ids.begin(), ids.end(),
                      [&](int32_t a, int32_t b) { return B.medianKey[axis][a] < B.medianKey[axis][b]; });
            const size_t half = ids.size() / 2;
            if (half == 1) L = B.wholeMesh(ids[0]);
            else L.meshIds.assign(ids.begin(), ids.begin() + half);
            if (ids.size() - half == 1) R = B.wholeMesh(ids[half]);
            else R.meshIds.assign(ids.begin() + half, ids.end());
        } else {
            // getChildMeshesOneMesh bvh.cpp:192-207 / sortTrianglesByCentres :122-134
            std::vector<int32_t>& p = B.perm[it.mesh];
            const std::vector<float>& k = B.keys[axis][it.mesh];
            std::sort(p.begin() + it.begin, p.begin() + it.end, [&](int32_t a, int32_t b) { return k[a] < k[b]; });
            const int32_t n = it.end - it.begin;
            L.mesh = R.mesh = it.mesh;
            L.begin = it.begin;
            L.end = it.begin + n / 2;
            R.begin = L.end;
            R.end = it.end;
        }
        const bool areLeaf = (nd.level + 1 == maxDepth - 1); // bvh.cpp:320
        const bool lLeaf = areLeaf || B.singleTriangle(L);
        const bool rLeaf = areLeaf || B.singleTriangle(R);
        const int32_t lastIndex = (int32_t)out.nodes.size();
        out.nodes[cur].child0 = lastIndex;
        out.nodes[cur].child1 = lastIndex + 1;
        items[cur] = Item();
        pushNode(std::move(L), nd.level + 1, lLeaf);
        pushNode(std::move(R), nd.level + 1, rLeaf);
    }

    int maxLevel = 0;
    for (size_t i = 0; i < out.nodes.size(); i++) {
        HostNode& n = out.nodes[i];
        if (n.level > maxLevel) maxLevel = n.level;
        if (!n.isLeaf) continue;
        n.firstTri = (int32_t)out.leafTris.size();
        B.forEachTriangle(items[i], [&](int32_t m, int32_t t) { out.leafTris.push_back(LeafTri{m, t}); });
        n.triCount = (int32_t)out.leafTris.size() - n.firstTri;
    }
    out.numLevels = maxLevel + 1; // numLevels() bvh.cpp:214-224
}

// =================================================================================================================
// Culling sub-trees inside the reference leaves
// =================================================================================================================
// Soundness argument (what "conservative" means here). The reference accepts triangle k for a ray iff, in fp32 and in the
// reference's evaluation order, the plane distance t passes its range checks and the point p = fl(o + d t) passes the three
// edge tests dot(n, cross(e_i, p - v_i)) >= 0 (src/ray_tracing.cpp:23-72). Each edge function is evaluated with an absolute
// error below ~2e-6 |e_i| |p - v_i|, so an accepted p lies inside the triangle grown by 2e-6 |p - v_i| per edge; for a
// triangle whose smallest angle is at least 0.01 rad this keeps p within 1e-3 * diameter of the triangle (thinner triangles get a
// proportionally larger margin, conservativeBox), and p itself is within 2^-22 * max|coordinate| of the exact ray. A sub-tree
// box is the union of its triangles' boxes grown by
//      eps = (1e-3 + growth(smallest angle)) * diameter(triangle) + 1e-6 * max|coordinate|
// (rounded outwards), hence every point the reference can accept lies inside the box of every ancestor of its triangle and
// the (tolerant) slab test of the traversal can never cull a triangle the reference would accept. Triangles that violate the
// angle condition or contain non-finite coordinates get an unbounded box.
namespace {

struct TriBox {
    float lo[3], hi[3], c[3];
    bool bounded;
};

inline float down(float x) { return std::nextafter(x, -std::numeric_limits<float>::infinity()); }
inline float up(float x) { return std::nextafter(x, std::numeric_limits<float>::infinity()); }

TriBox conservativeBox(const float* p0, const float* p1, const float* p2)
{
    TriBox tb;
    bool finite = true;
    double q[3][3];
    for (int k = 0; k < 3; k++) {
        q[0][k] = p0[k]; q[1][k] = p1[k]; q[2][k] = p2[k];
        finite = finite && std::isfinite(p0[k]) && std::isfinite(p1[k]) && std::isfinite(p2[k]);
    }
    bool bounded = finite;
    double diam = 0.0, maxAbs = 0.0, grow = 0.0;
    if (finite) {
        double e[3][3], len[3];
        for (int i = 0; i < 3; i++) {
            len[i] = 0.0;
            for (int k = 0; k < 3; k++) {
                e[i][k] = q[(i + 1) % 3][k] - q[i][k];
                len[i] += e[i][k] * e[i][k];
                maxAbs = std::max(maxAbs, std::fabs(q[i][k]));
            }
            len[i] = std::sqrt(len[i]);
            diam = std::max(diam, len[i]);
        }
        // smallest angle via the largest edge: sin(angle) = 2*area / (a*b)
        const double cx = e[0][1] * e[2][2] - e[0][2] * e[2][1], cy = e[0][2] * e[2][0] - e[0][0] * e[2][2],
                     cz = e[0][0] * e[2][1] - e[0][1] * e[2][0];
        const double area2 = std::sqrt(cx * cx + cy * cy + cz * cz);
        double minSin = 1.0;
        for (int i = 0; i < 3; i++) {
            const double prod = len[i] * len[(i + 2) % 3]; // the two edges meeting at vertex i
            if (!(prod > 0.0)) { minSin = 0.0; break; }
            minSin = std::min(minSin, area2 / prod);
        }
        // growth of the accept region: offsetting the two edges at a vertex of angle theta by delta = 2e-6 |p - v| moves the
        // corner by delta / sin(theta/2) ~ 4e-6 R / sin(theta), R = extent of the region: R <= diam / (1 - g), g = 4e-6 / minSin.
        // The computed plane normal of a sliver is tilted by ~6e-8 / sin(theta) (cancellation in the cross product), which
        // moves the hit point by less than that fraction of diam. Both stay far below the margin while minSin >= 2e-5;
        // thinner, degenerate (cross product below the normal range) or huge triangles cannot be bounded: always tested.
        if (!(minSin >= 2e-5) || !(diam < 1e30) || !(area2 >= 1e-30)) bounded = false;
        else grow = 2.0 * (4e-6 / minSin) / (1.0 - 4e-6 / minSin);
    }
    tb.bounded = bounded;
    for (int k = 0; k < 3; k++) {
        if (!bounded) {
            tb.lo[k] = -FLT_MAX;
            tb.hi[k] = FLT_MAX;
            tb.c[k] = finite ? (float)((q[0][k] + q[1][k] + q[2][k]) / 3.0) : 0.0f;
            continue;
        }
        const double eps = (1e-3 + grow) * diam + 1e-6 * maxAbs;
        const double lo = std::min(q[0][k], std::min(q[1][k], q[2][k])) - eps;
        const double hi = std::max(q[0][k], std::max(q[1][k], q[2][k])) + eps;
        tb.lo[k] = down((float)lo);
        tb.hi[k] = up((float)hi);
        tb.c[k] = (float)((q[0][k] + q[1][k] + q[2][k]) / 3.0);
    }
    return tb;
}

struct WideBuilder {
    const std::vector<TriBox>& boxes; // per position (leaf order at entry)
    std::vector<int32_t>& order;      // positions being permuted (values index `boxes`)
    std::vector<WideNode>& nodes;
    int subLeafSize;
    int base; // first position of the reference leaf in the global arrays

    void bounds(int begin, int end, float lo[3], float hi[3]) const
    {
        for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
        for (int i = begin; i < end; i++) {
            const TriBox& b = boxes[order[i]];
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], b.lo[k]);
                hi[k] = std::max(hi[k], b.hi[k]);
            }
        }
    }

    // median split of [begin,end) along the longest centroid axis; returns the split point
    int split(int begin, int end)
    {
        float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = begin; i < end; i++)
            for (int k = 0; k < 3; k++) {
                clo[k] = std::min(clo[k], boxes[order[i]].c[k]);
                chi[k] = std::max(chi[k], boxes[order[i]].c[k]);
            }
        int axis = 0;
        if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
        if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
        const int mid = begin + (end - begin) / 2;
        std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end, [&](int32_t x, int32_t y) {
            const float cx = boxes[x].c[axis], cy = boxes[y].c[axis];
            return cx < cy || (cx == cy && x < y);
        });
        return mid;
    }

    // builds the wide node for [begin,end) into nodes[self]
    void build(int self, int begin, int end)
    {
        // three binary levels collapsed: up to 8 sub-ranges, a range is not split further once it fits a sub-leaf
        std::vector<std::pair<int, int>> ranges{{begin, end}};
        for (int level = 0; level < 3; level++) {
            std::vector<std::pair<int, int>> next;
            for (const auto& r : ranges) {
                if (r.second - r.first <= subLeafSize) {
                    next.push_back(r);
                    continue;
                }
                const int mid = split(r.first, r.second);
                next.push_back({r.first, mid});
                next.push_back({mid, r.second});
            }
            ranges.swap(next);
        }
        WideNode n;
        for (int c = 0; c < 8; c++) {
            for (int k = 0; k < 3; k++) { n.lo[c][k] = FLT_MAX; n.hi[c][k] = -FLT_MAX; } // inverted: never hit
            n.id[c] = 0u;
        }
        std::vector<int> childNode(ranges.size(), -1);
        for (size_t c = 0; c < ranges.size(); c++) {
            const int b = ranges[c].first, e = ranges[c].second;
            bounds(b, e, n.lo[c], n.hi[c]);
            if (e - b <= subLeafSize) {
                n.id[c] = 0x40000000u | 0x20000000u | ((uint32_t)(e - b - 1) << 26) | (uint32_t)(base + b);
            } else {
                childNode[c] = (int)nodes.size();
                nodes.push_back(WideNode());
                n.id[c] = 0x40000000u | (uint32_t)childNode[c];
            }
        }
        nodes[self] = n;
        for (size_t c = 0; c < ranges.size(); c++)
            if (childNode[c] >= 0) build(childNode[c], ranges[c].first, ranges[c].second);
    }
};

} // namespace

void buildLeafSubTrees(const std::vector<MeshView>& meshes, BuiltBVH& bvh, int minLeafForSubTree, int subLeafSize)
{
    if (subLeafSize > 8) subLeafSize = 8;
    if (subLeafSize < 1) subLeafSize = 1;
    const size_t T = bvh.leafTris.size();
    bvh.leafTrisReferenceOrder = bvh.leafTris;
    bvh.leafRank.assign(T, 0);
    bvh.wide.clear();
    bvh.wideRoot.assign(bvh.nodes.size(), -1);
    std::vector<LeafTri> permuted = bvh.leafTris;
    for (size_t ni = 0; ni < bvh.nodes.size(); ni++) {
        const HostNode& n = bvh.nodes[ni];
        if (!n.isLeaf) continue;
        const int first = n.firstTri, count = n.triCount;
        for (int i = 0; i < count; i++) bvh.leafRank[first + i] = i;
        if (count < minLeafForSubTree || count <= subLeafSize) continue;
        std::vector<TriBox> boxes(count);
        for (int i = 0; i < count; i++) {
            const LeafTri lt = bvh.leafTris[first + i];
            const MeshView& mv = meshes[lt.mesh];
            const uint32_t* tri = mv.triangles + 3 * (size_t)lt.tri;
            boxes[i] = conservativeBox(mv.vertices + 6 * (size_t)tri[0], mv.vertices + 6 * (size_t)tri[1],
                                       mv.vertices + 6 * (size_t)tri[2]);
        }
        std::vector<int32_t> order(count);
        for (int i = 0; i < count; i++) order[i] = i;
        const int root = (int)bvh.wide.size();
        bvh.wide.push_back(WideNode());
        WideBuilder wb{boxes, order, bvh.wide, subLeafSize, first};
        wb.build(root, 0, count);
        bvh.wideRoot[ni] = root;
        for (int i = 0; i < count; i++) {
            permuted[first + i] = bvh.leafTris[first + order[i]];
            bvh.leafRank[first + i] = order[i];
        }
    }
    bvh.leafTris.swap(permuted);
}

// =================================================================================================================
// Fast tree: the reference tree collapsed into 8-wide conservative nodes on top of the leaf sub-trees
// =================================================================================================================
namespace {

// ---- binned SAH over conservative triangle boxes -> binary tree -> 8-wide nodes ---------------------------------------------
struct SahNode {
    float lo[3], hi[3];
    int left, right;  // -1 for leaves
    int first, count; // leaves: range of the order array
};

#ifndef SAH_BINS
#define SAH_BINS 16
#endif
struct SahBuilder {
    const std::vector<TriBox>& boxes; // per position
    std::vector<int32_t>& order;      // positions, permuted in place; leaves are contiguous ranges
    std::vector<SahNode> nodes;
    int leafMax;

    static double area(const float lo[3], const float hi[3])
    {
        const double dx = std::max(0.0, (double)hi[0] - lo[0]), dy = std::max(0.0, (double)hi[1] - lo[1]),
                     dz = std::max(0.0, (double)hi[2] - lo[2]);
        return dx * dy + dy * dz + dz * dx;
    }

    int build(int begin, int end)
    {
        const int self = (int)nodes.size();
        nodes.push_back(SahNode());
        SahNode n;
        float clo[3], chi[3];
        for (int k = 0; k < 3; k++) { n.lo[k] = clo[k] = FLT_MAX; n.hi[k] = chi[k] = -FLT_MAX; }
        for (int i = begin; i < end; i++) {
            const TriBox& b = boxes[order[i]];
            for (int k = 0; k < 3; k++) {
                n.lo[k] = std::min(n.lo[k], b.lo[k]); n.hi[k] = std::max(n.hi[k], b.hi[k]);
                clo[k] = std::min(clo[k], b.c[k]); chi[k] = std::max(chi[k], b.c[k]);
            }
        }
        n.left = n.right = -1;
        n.first = begin;
        n.count = end - begin;
        const int N = end - begin;
        int mid = -1;
        if (N > leafMax) {
            // 16 bins per axis over the centroid bounds; cost of a split = A_L N_L + A_R N_R
            const int NB = SAH_BINS;
            double bestCost = 1e300;
            int bestAxis = -1, bestBin = -1;
            for (int axis = 0; axis < 3; axis++) {
                const float ext = chi[axis] - clo[axis];
                if (!(ext > 0.0f) || !std::isfinite(ext)) continue;
                int cnt[NB] = {0};
                float blo[NB][3], bhi[NB][3];
                for (int b = 0; b < NB; b++)
                    for (int k = 0; k < 3; k++) { blo[b][k] = FLT_MAX; bhi[b][k] = -FLT_MAX; }
                const float scale = NB / ext;
                for (int i = begin; i < end; i++) {
                    const TriBox& t = boxes[order[i]];
                    int b = (int)((t.c[axis] - clo[axis]) * scale);
                    b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                    cnt[b]++;
                    for (int k = 0; k < 3; k++) { blo[b][k] = std::min(blo[b][k], t.lo[k]); bhi[b][k] = std::max(bhi[b][k], t.hi[k]); }
                }
                double rightA[NB];
                int rightN[NB];
                float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                int acc = 0;
                for (int b = NB - 1; b >= 1; b--) {
                    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], blo[b][k]); hi[k] = std::max(hi[k], bhi[b][k]); }
                    acc += cnt[b];
                    rightA[b] = area(lo, hi);
                    rightN[b] = acc;
                }
                for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
                acc = 0;
                for (int b = 0; b < NB - 1; b++) { // split between bin b and b + 1
                    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], blo[b][k]); hi[k] = std::max(hi[k], bhi[b][k]); }
                    acc += cnt[b];
                    if (acc == 0 || rightN[b + 1] == 0) continue;
                    const double cost = area(lo, hi) * acc + rightA[b + 1] * rightN[b + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
                }
            }
            const bool worthIt = bestAxis >= 0 && (N > 8 || bestCost < area(n.lo, n.hi) * N);
            if (worthIt) {
                const float ext = chi[bestAxis] - clo[bestAxis];
                const float scale = NB / ext;
                auto it = std::partition(order.begin() + begin, order.begin() + end, [&](int32_t p) {
                    int b = (int)((boxes[p].c[bestAxis] - clo[bestAxis]) * scale);
                    b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                    return b <= bestBin;
                });
                mid = (int)(it - order.begin());
                if (mid == begin || mid == end) mid = -1;
            }
            if (mid < 0 && N > 8) { // no usable plane (coincident centroids ...): split the range in two halves
                mid = begin + N / 2;
            }
        }
        if (mid >= 0) {
            n.left = build(begin, mid);
            n.right = build(mid, end);
        }
        nodes[self] = n;
        return self;
    }
};

} // namespace

static uint32_t buildFastTreeSah(const std::vector<TriBox>& boxes, std::vector<int32_t>& order, BuiltBVH& bvh)
{
    const uint32_t ID_TRI = 0x20000000u, ID_SUB = 0x40000000u;
    if (order.empty()) return 0u;
    // The binary SAH tree is built down to 2-triangle leaves; the 8-wide collapse decides the real leaf size: a slot whose
    // subtree holds at most `leafCap` triangles becomes a triangle leaf, and while a node has free slots the largest slots keep
    // being opened (also below leafCap): all 8 box tests of a step are paid anyway, fuller nodes mean fewer triangle tests.
    int leafCap = 6; // triangles per leaf (speed only; CGRT_SAH_LEAF overrides for tuning)
    if (const char* e = getenv("CGRT_SAH_LEAF")) leafCap = std::max(1, std::min(8, atoi(e)));
    int binLeaf = leafCap; // default: binary leaves = wide leaves (the configuration measured on the GPU, profiles/r01_tuning.md);
                           // CGRT_SAH_BINLEAF=2 gives 27 % fewer triangle tests for 17 % more leaf steps on the CPU proxy
                           // (tests/tree_work.py) - to be measured on the GPU
    if (const char* e = getenv("CGRT_SAH_BINLEAF")) binLeaf = std::max(1, std::min(8, atoi(e)));
    int minOpen = 2; // slots with fewer triangles than this are not opened further
    if (const char* e = getenv("CGRT_SAH_MINOPEN")) minOpen = std::max(2, atoi(e));
    SahBuilder sb{boxes, order, {}, std::min(binLeaf, leafCap)};
    sb.nodes.reserve(order.size());
    const int root = sb.build(0, (int)order.size());
    std::function<uint32_t(int)> wide = [&](int bi) -> uint32_t {
        const SahNode& b = sb.nodes[bi];
        if (b.left < 0 || b.count <= leafCap) return ID_SUB | ID_TRI | ((uint32_t)(b.count - 1) << 26) | (uint32_t)b.first;
        std::vector<int> slots{b.left, b.right};
        while (slots.size() < 8) {
            int best = -1;
            double bestA = -1.0;
            for (size_t c = 0; c < slots.size(); c++) {
                const SahNode& sn = sb.nodes[slots[c]];
                if (sn.left < 0 || sn.count < minOpen) continue;
                const double a = SahBuilder::area(sn.lo, sn.hi);
                if (a > bestA) { bestA = a; best = (int)c; }
            }
            if (best < 0) break;
            const SahNode o = sb.nodes[slots[best]];
            slots[best] = o.left;
            slots.push_back(o.right);
        }
        const int self = (int)bvh.wide.size();
        bvh.wide.push_back(WideNode());
        WideNode w;
        for (int c = 0; c < 8; c++) {
            for (int k = 0; k < 3; k++) { w.lo[c][k] = FLT_MAX; w.hi[c][k] = -FLT_MAX; }
            w.id[c] = 0u;
        }
        for (size_t c = 0; c < slots.size(); c++) {
            const SahNode& sn = sb.nodes[slots[c]];
            for (int k = 0; k < 3; k++) { w.lo[c][k] = sn.lo[k]; w.hi[c][k] = sn.hi[k]; }
            w.id[c] = wide(slots[c]);
        }
        bvh.wide[self] = w;
        return ID_SUB | (uint32_t)self;
    };
    return wide(root);
}

void buildFastTree(const std::vector<MeshView>& meshes, BuiltBVH& bvh, bool sah)
{
    const size_t NN = bvh.nodes.size();
    bvh.fastRoot = 0u;
    bvh.fastOrder.clear();
    bvh.alwaysTest.clear();
    bvh.parent.assign(NN, -1);
    bvh.triLeafNode.assign(bvh.leafTris.size(), 0);
    if (NN == 0) return;
    if (bvh.wideRoot.size() != NN) bvh.wideRoot.assign(NN, -1);
    // conservative box of every reference node, bottom-up (children always have larger indices: BFS numbering)
    std::vector<TriBox> cons(NN);
    std::vector<TriBox> triBoxes(sah ? bvh.leafTris.size() : 0);
    for (size_t i = NN; i-- > 0;) {
        const HostNode& n = bvh.nodes[i];
        TriBox& b = cons[i];
        for (int k = 0; k < 3; k++) { b.lo[k] = FLT_MAX; b.hi[k] = -FLT_MAX; b.c[k] = 0.0f; }
        if (n.isLeaf) {
            for (int t = 0; t < n.triCount; t++) {
                const LeafTri lt = bvh.leafTris[n.firstTri + t];
                bvh.triLeafNode[n.firstTri + t] = (int32_t)i;
                const MeshView& mv = meshes[lt.mesh];
                const uint32_t* tri = mv.triangles + 3 * (size_t)lt.tri;
                const TriBox tb = conservativeBox(mv.vertices + 6 * (size_t)tri[0], mv.vertices + 6 * (size_t)tri[1],
                                                  mv.vertices + 6 * (size_t)tri[2]);
                if (sah) triBoxes[n.firstTri + t] = tb;
                if (!tb.bounded) { // would make every ancestor's box unbounded: kept out of the tree, tested for every ray
                    bvh.alwaysTest.push_back(n.firstTri + t);
                    continue;
                }
                for (int k = 0; k < 3; k++) { b.lo[k] = std::min(b.lo[k], tb.lo[k]); b.hi[k] = std::max(b.hi[k], tb.hi[k]); }
            }
        } else {
            bvh.parent[n.child0] = (int32_t)i;
            bvh.parent[n.child1] = (int32_t)i;
            for (int c = 0; c < 2; c++) {
                const TriBox& cb = cons[c == 0 ? n.child0 : n.child1];
                for (int k = 0; k < 3; k++) { b.lo[k] = std::min(b.lo[k], cb.lo[k]); b.hi[k] = std::max(b.hi[k], cb.hi[k]); }
            }
        }
    }
    if (sah) { // independent tree with its own triangle order; unbounded triangles go last (reached only through alwaysTest)
        std::vector<char> unb(bvh.leafTris.size(), 0);
        for (int32_t a : bvh.alwaysTest) unb[a] = 1;
        std::vector<int32_t> order;
        order.reserve(bvh.leafTris.size());
        for (size_t p = 0; p < bvh.leafTris.size(); p++)
            if (!unb[p]) order.push_back((int32_t)p);
        const uint32_t root = bvh.alwaysTest.size() <= 256 ? buildFastTreeSah(triBoxes, order, bvh) : 0u;
        bvh.fastOrder = order;
        for (int32_t a : bvh.alwaysTest) bvh.fastOrder.push_back(a);
        bvh.fastRoot = root;
        return;
    }
    const uint32_t ID_TRI = 0x20000000u, ID_SUB = 0x40000000u;
    // id of a reference leaf in the fast tree: its sub-tree root, or its triangles directly (leaves without a sub-tree hold at
    // most 7 triangles, buildLeafSubTrees), or nothing when it is empty. 0xffffffff = cannot be represented.
    auto leafId = [&](int ni) -> uint32_t {
        const HostNode& n = bvh.nodes[ni];
        if (bvh.wideRoot[ni] >= 0) return ID_SUB | (uint32_t)bvh.wideRoot[ni];
        if (n.triCount <= 0) return 0u;
        if (n.triCount > 8) return 0xffffffffu;
        return ID_SUB | ID_TRI | ((uint32_t)(n.triCount - 1) << 26) | (uint32_t)n.firstTri;
    };
    auto area = [&](int ni) -> double {
        const TriBox& b = cons[ni];
        const double dx = std::max(0.0, (double)b.hi[0] - b.lo[0]), dy = std::max(0.0, (double)b.hi[1] - b.lo[1]),
                     dz = std::max(0.0, (double)b.hi[2] - b.lo[2]);
        const double a = dx * dy + dy * dz + dz * dx;
        return std::isfinite(a) ? a : 1e300;
    };
    bool ok = true;
    // wide node for reference inner node `ni`: open the child with the largest box until eight slots are used
    std::function<uint32_t(int)> build = [&](int ni) -> uint32_t {
        const HostNode& n = bvh.nodes[ni];
        if (n.isLeaf) {
            const uint32_t id = leafId(ni);
            if (id == 0xffffffffu) ok = false;
            return id;
        }
        std::vector<int> slots{n.child0, n.child1};
        while (slots.size() < 8) {
            int best = -1;
            double bestA = -1.0;
            for (size_t c = 0; c < slots.size(); c++)
                if (!bvh.nodes[slots[c]].isLeaf && area(slots[c]) > bestA) { bestA = area(slots[c]); best = (int)c; }
            if (best < 0) break;
            const HostNode& o = bvh.nodes[slots[best]];
            slots[best] = o.child0;
            slots.push_back(o.child1);
        }
        const int self = (int)bvh.wide.size();
        bvh.wide.push_back(WideNode());
        WideNode w;
        for (int c = 0; c < 8; c++) {
            for (int k = 0; k < 3; k++) { w.lo[c][k] = FLT_MAX; w.hi[c][k] = -FLT_MAX; }
            w.id[c] = 0u;
        }
        for (size_t c = 0; c < slots.size(); c++) {
            const uint32_t id = build(slots[c]);
            if (id == 0u || id == 0xffffffffu) continue; // empty leaf: slot stays unused
            for (int k = 0; k < 3; k++) { w.lo[c][k] = cons[slots[c]].lo[k]; w.hi[c][k] = cons[slots[c]].hi[k]; }
            w.id[c] = id;
        }
        bvh.wide[self] = w;
        return ID_SUB | (uint32_t)self;
    };
    const uint32_t root = build(0);
    // a scene with many unboundable triangles is not worth a speculative search (each costs every ray one exact test)
    bvh.fastRoot = (ok && root != 0xffffffffu && bvh.alwaysTest.size() <= 256) ? root : 0u;
}

} // namespace cgrt
