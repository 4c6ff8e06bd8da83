// TEST INFRASTRUCTURE (CPU): compiles the PRODUCT's traversal source (cg-raytracer_b200/csrc/cgrt_device.cuh, rt_math.cuh) for
// the host and runs, ray by ray, (i) traverseStrict - the literal reference-order traversal with sequential leaf scans - and
// (ii) the speculative search + certificate (fastStart / fastStep / fastFinish) WITHOUT the exact replay. Whenever the
// certificate accepts, the two results must be identical in every bit; tests/test_spec_certificate_cpu.py fuzzes that with
// millions of rays on adversarial scenes (axis-aligned walls in box faces, coplanar duplicates, rays through shared edges and
// vertices, zero direction components, slivers). Same arithmetic as the GPU build: IEEE single, no contraction (-ffp-contract=off).
// The scene flattening below restates what cgrt_capi.cu uploads (nodes, tri4, tri4f, wide, refParent, alwaysTri).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }

#include "../../cg-raytracer_b200/csrc/cgrt_device.cuh"
#include "../../cg-raytracer_b200/csrc/bvh_build.h"
#include "../../include/cgrt_b200.h"

using namespace cgrt;

namespace {
struct HostScene {
    BuiltBVH bvh;
    std::vector<float4> nodes, triPl, v[3], n[3], tri4, tri4f, wide, pairs;
    std::vector<int> parent, always;
    int rootId = 0;
    DevScene S;
};

void flatten(const cgrt_scene_desc* d, int maxDepth, bool sah, HostScene& H)
{
    std::vector<MeshView> views;
    size_t vo = 0, to = 0;
    int32_t gid = 0;
    for (int m = 0; m < d->n_meshes; m++) {
        MeshView mv;
        mv.vertices = d->vertices + 6 * vo;
        mv.triangles = d->triangles + 3 * to;
        mv.nv = d->mesh_vertex_count[m];
        mv.nt = d->mesh_triangle_count[m];
        mv.triOffset = gid;
        views.push_back(mv);
        vo += mv.nv; to += mv.nt; gid += mv.nt;
    }
    buildReferenceBVH(views, maxDepth, H.bvh);
    buildLeafSubTrees(views, H.bvh);
    buildFastTree(views, H.bvh, sah);
    const size_t T = H.bvh.leafTris.size(), NN = H.bvh.nodes.size();
    H.nodes.resize(NN * 2);
    for (size_t i = 0; i < NN; i++) {
        const HostNode& n = H.bvh.nodes[i];
        const uint32_t a = n.isLeaf ? (uint32_t)n.firstTri : (uint32_t)n.child0, b = n.isLeaf ? (uint32_t)n.triCount : 0u;
        H.nodes[2 * i] = make_float4(n.lo[0], n.lo[1], n.lo[2], __int_as_float((int)a));
        H.nodes[2 * i + 1] = make_float4(n.hi[0], n.hi[1], n.hi[2], __int_as_float((int)b));
    }
    for (int k = 0; k < 3; k++) { H.v[k].resize(T); H.n[k].resize(T); }
    H.triPl.resize(T); H.tri4.resize(4 * T); H.tri4f.resize(4 * T);
    for (size_t i = 0; i < T; i++) {
        const LeafTri lt = H.bvh.leafTris[i];
        const MeshView& mv = views[lt.mesh];
        const int32_t g = mv.triOffset + lt.tri;
        for (int k = 0; k < 3; k++) {
            const float* vtx = mv.vertices + 6 * (size_t)mv.triangles[3 * (size_t)lt.tri + k];
            const int w = k == 0 ? g : (k == 1 ? lt.mesh : H.bvh.leafRank[i]);
            H.v[k][i] = make_float4(vtx[0], vtx[1], vtx[2], __int_as_float(w));
            H.n[k][i] = make_float4(vtx[3], vtx[4], vtx[5], k == 0 ? __int_as_float(H.bvh.triLeafNode.size() > i ? H.bvh.triLeafNode[i] : 0) : 0.0f);
        }
        H.triPl[i] = trianglePlaneDev(mk3(H.v[0][i]), mk3(H.v[1][i]), mk3(H.v[2][i])); // same expression tree as k_setup_planes
        H.tri4[4 * i + 0] = H.triPl[i];
        H.tri4[4 * i + 1] = H.v[0][i];
        H.tri4[4 * i + 2] = H.v[1][i];
        H.tri4[4 * i + 3] = H.v[2][i];
    }
    for (size_t k = 0; k < T; k++) { // k_permute_tri4
        const size_t p = H.bvh.fastOrder.size() == T ? (size_t)H.bvh.fastOrder[k] : k;
        for (int q = 0; q < 4; q++) H.tri4f[4 * k + q] = H.tri4[4 * p + q];
        H.tri4f[4 * k + 3].w = __int_as_float((int)p);
    }
    H.wide.assign(H.bvh.wide.size() * 16, make_float4(0, 0, 0, 0));
    for (size_t j = 0; j < H.bvh.wide.size(); j++) {
        const WideNode& n = H.bvh.wide[j];
        float4* out = H.wide.data() + 16 * j;
        for (int k = 0; k < 3; k++)
            for (int h = 0; h < 2; h++) {
                out[2 * k + h] = make_float4(n.lo[4 * h][k], n.lo[4 * h + 1][k], n.lo[4 * h + 2][k], n.lo[4 * h + 3][k]);
                out[6 + 2 * k + h] = make_float4(n.hi[4 * h][k], n.hi[4 * h + 1][k], n.hi[4 * h + 2][k], n.hi[4 * h + 3][k]);
            }
        for (int h = 0; h < 2; h++) {
            float f[4];
            std::memcpy(f, &n.id[4 * h], 16);
            out[12 + h] = make_float4(f[0], f[1], f[2], f[3]);
        }
    }
    // the exact traversal's child-pair array (cgrt_capi.cu): both children of every inner reference node with their visit ids
    {
        const uint32_t ID_REFLEAF = 0x10000000u, ID_REFSCAN = 0x80000000u;
        auto refId = [&](int cidx) -> uint32_t {
            const HostNode& n = H.bvh.nodes[cidx];
            if (!n.isLeaf) return (uint32_t)((n.child0 - 1) / 2);
            const int wr = H.bvh.wideRoot[cidx];
            if (wr >= 0) return ID_REFLEAF | (uint32_t)wr;
            return ID_REFSCAN | (uint32_t)cidx;
        };
        const int nPairs = NN > 0 ? (int)(NN - 1) / 2 : 0;
        H.pairs.assign((size_t)nPairs * 4, make_float4(0, 0, 0, 0));
        for (size_t i = 0; i < NN; i++) {
            const HostNode& n = H.bvh.nodes[i];
            if (n.isLeaf) continue;
            float4* out = H.pairs.data() + 4 * (size_t)((n.child0 - 1) / 2);
            const HostNode &l = H.bvh.nodes[n.child0], &r = H.bvh.nodes[n.child1];
            out[0] = make_float4(l.lo[0], l.lo[1], l.lo[2], __int_as_float((int)refId(n.child0)));
            out[1] = make_float4(l.hi[0], l.hi[1], l.hi[2], __int_as_float(n.child0));
            out[2] = make_float4(r.lo[0], r.lo[1], r.lo[2], __int_as_float((int)refId(n.child1)));
            out[3] = make_float4(r.hi[0], r.hi[1], r.hi[2], __int_as_float(n.child1));
        }
        H.S.rootId = 0;
        H.rootId = NN > 0 ? (int)refId(0) : 0;
    }
    H.parent.assign(H.bvh.parent.begin(), H.bvh.parent.end());
    if (H.parent.empty()) H.parent.assign(NN ? NN : 1, -1);
    H.always.assign(H.bvh.alwaysTest.begin(), H.bvh.alwaysTest.end());
    std::memset(&H.S, 0, sizeof H.S);
    H.S.nodes = H.nodes.data();
    H.S.triPl = H.triPl.data();
    H.S.tri4 = H.tri4.data();
    H.S.tri4f = H.tri4f.data();
    H.S.triV0 = H.v[0].data(); H.S.triV1 = H.v[1].data(); H.S.triV2 = H.v[2].data();
    H.S.triN0 = H.n[0].data(); H.S.triN1 = H.n[1].data(); H.S.triN2 = H.n[2].data();
    H.S.wide = H.wide.data();
    H.S.pairs = H.pairs.data();
    H.S.rootId = H.rootId;
    H.S.refParent = H.parent.data();
    H.S.alwaysTri = H.always.data();
    H.S.nAlways = H.bvh.fastRoot != 0u ? (int)H.always.size() : 0;
    H.S.fastRoot = H.bvh.fastRoot;
    H.S.nNodes = (int)NN;
    H.S.nTris = (int)T;
    H.S.nSpheres = 0;
    H.S.nMeshes = d->n_meshes;
}
} // namespace

extern "C" {

// search work of the speculative traversal for a ray set (closest hit): totals of 8-wide node steps, leaf steps and triangle
// tests - a CPU proxy for tuning the fast tree (builder flavour, leaf size) without a GPU. out[4] = wide, leaf, triangles, max steps
int spec_work(const cgrt_scene_desc* d, int max_depth, int sah, const float* rays, int64_t n, int64_t* out)
{
    HostScene H;
    flatten(d, max_depth > 0 ? max_depth : 12, sah != 0, H);
    const DevScene& S = H.S;
    int64_t nw = 0, nl = 0, ntri = 0, mx = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : nw, nl, ntri) reduction(max : mx)
    for (int64_t i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        FastTrav T;
        FastStack K;
        int state = fastStart<false>(S, T, K, mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], 0.0f, 0.0f);
        int64_t steps = 0;
        while (state == TRAV_CONTINUE) {
            if (travIsLeaf(T.node)) { nl++; ntri += (int64_t)((T.node >> CGRT_TRICNT_SHIFT) & 7u) + 1; }
            else nw++;
            steps++;
            state = fastStep<false>(S, T, K, 0.0f, 0.0f);
        }
        if (steps > mx) mx = steps;
    }
    out[0] = nw; out[1] = nl; out[2] = ntri; out[3] = mx;
    return 0;
}

// rays: [n][8] = origin, t, direction, pad. mode 0: closest hit, mode 1: any hit (max_dist[n], eps); modes 2 / 3: the same
// two queries through the EXACT replay traversal (traverseFast), which must equal the literal traversal for every ray.
// out_exact / out_fast: [n][2] int32 = (global triangle id or -1 | shadowed flag, t bits | 0); certified[n] = 1 when the speculative
// result carries a certificate (or is a certain miss). stats[8] = rays, certified, deferred, mismatches, first mismatch,
// fast tree present, always-list size, wide nodes.
int spec_run(const cgrt_scene_desc* d, int max_depth, int sah, const float* rays, int64_t n, int mode, const float* max_dist,
             float eps, int32_t* out_exact, int32_t* out_fast, uint8_t* certified, int64_t* stats)
{
    HostScene H;
    flatten(d, max_depth > 0 ? max_depth : 12, sah != 0, H);
    const DevScene& S = H.S;
    int64_t nCert = 0, nDefer = 0, nBad = 0, firstBad = -1;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : nCert, nDefer, nBad)
    for (int64_t i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        const V3 o = mk3(r[0], r[1], r[2]), dd = mk3(r[4], r[5], r[6]);
        const float tIn = r[3];
        TraceResult E;
        uint32_t nb = 0, nt = 0;
        const bool hitE = traverseStrict<false, false>(S, o, dd, tIn, 0.0f, 0.0f, E, nb, nt);
        TraceResult F;
        F.sphere = -1; F.tri = -1; F.t = tIn;
        bool defer = false, resF;
        FastTrav T;
        FastStack K;
        if (mode == 2 || mode == 3) { // the exact replay traversal (filtered reference decisions + leaf sub-trees) vs the literal one
            const float md = mode == 3 ? max_dist[i] : 0.0f;
            TraceResult X;
            const bool r = mode == 3 ? traverseFast<true>(S, o, dd, tIn, eps, md, X) : traverseFast<false>(S, o, dd, tIn, 0.0f, 0.0f, X);
            if (mode == 2) {
                auto gidOf = [&](int pos) { return pos >= 0 ? __float_as_int(H.v[0][pos].w) : -1; };
                out_exact[2 * i] = hitE ? gidOf(E.tri) : -1;
                out_exact[2 * i + 1] = __float_as_int(hitE ? E.t : tIn);
                out_fast[2 * i] = r ? gidOf(X.tri) : -1;
                out_fast[2 * i + 1] = __float_as_int(r ? X.t : tIn);
            } else {
                out_exact[2 * i] = (hitE && !(E.t + eps >= md)) ? 1 : 0;
                out_fast[2 * i] = r ? 1 : 0;
                out_exact[2 * i + 1] = out_fast[2 * i + 1] = 0;
            }
            if (out_exact[2 * i] != out_fast[2 * i] || out_exact[2 * i + 1] != out_fast[2 * i + 1]) {
                nBad++;
#pragma omp critical
                if (firstBad < 0 || i < firstBad) firstBad = i;
            }
            certified[i] = 1;
            nCert++;
            continue;
        }
        if (mode == 0) {
            int state = fastStart<false>(S, T, K, o, dd, tIn, 0.0f, 0.0f);
            while (state == TRAV_CONTINUE) state = fastStep<false>(S, T, K, 0.0f, 0.0f);
            resF = fastFinish<false>(S, T, K.t2, state, tIn, 0.0f, 0.0f, F, defer);
            auto gidOf = [&](int pos) { return pos >= 0 ? __float_as_int(H.v[0][pos].w) : -1; }; // global triangle id
            out_exact[2 * i] = hitE ? gidOf(E.tri) : -1;
            out_exact[2 * i + 1] = __float_as_int(hitE ? E.t : tIn);
            out_fast[2 * i] = (!defer && resF) ? gidOf(F.tri) : -1;
            out_fast[2 * i + 1] = __float_as_int((!defer && resF) ? F.t : tIn);
            if (!defer && (out_exact[2 * i] != out_fast[2 * i] || out_exact[2 * i + 1] != out_fast[2 * i + 1])) {
                nBad++;
#pragma omp critical
                if (firstBad < 0 || i < firstBad) firstBad = i;
            }
        } else {
            const float md = max_dist[i];
            const bool shE = hitE && !(E.t + eps >= md); // pointInShadow, src/main.cpp:115-131
            int state = fastStart<true>(S, T, K, o, dd, tIn, eps, md);
            while (state == TRAV_CONTINUE) state = fastStep<true>(S, T, K, eps, md);
            resF = fastFinish<true>(S, T, K.t2, state, tIn, eps, md, F, defer);
            out_exact[2 * i] = shE ? 1 : 0;
            out_exact[2 * i + 1] = 0;
            out_fast[2 * i] = (!defer && resF) ? 1 : 0;
            out_fast[2 * i + 1] = 0;
            if (!defer && out_exact[2 * i] != out_fast[2 * i]) {
                nBad++;
#pragma omp critical
                if (firstBad < 0 || i < firstBad) firstBad = i;
            }
        }
        certified[i] = defer ? 0 : 1;
        if (defer) nDefer++; else nCert++;
    }
    stats[0] = n; stats[1] = nCert; stats[2] = nDefer; stats[3] = nBad; stats[4] = firstBad;
    stats[5] = H.bvh.fastRoot != 0u; stats[6] = (int64_t)H.always.size(); stats[7] = (int64_t)H.bvh.wide.size();
    return 0;
}

// Property check of the certificates' first stage (cgrt_device.cuh slabFastHit): whenever it claims a hit, the reference's own
// slabTest (six IEEE divisions) must report a hit for an unbounded ray, at a distance no larger than the claimed upper bound.
// Inputs: n random (box, ray) pairs from a generator that favours the hard cases - origins on / next to faces, grazing rays,
// tiny and huge direction components, flat boxes, far origins. out[4] = cases, fast-stage claims, violations, fast-stage
// claims whose bound is more than 1e-5 (relative) above the reference's distance (looseness, not an error).
int spec_slab_fast_check(int64_t n, uint64_t seed, int64_t* out)
{
    int64_t claims = 0, bad = 0, loose = 0;
#pragma omp parallel for schedule(static) reduction(+ : claims, bad, loose)
    for (int64_t i = 0; i < n; i++) {
        uint64_t x = seed * 0x9E3779B97F4A7C15ull + (uint64_t)i * 0xD1B54A32D192ED03ull + 1;
        auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
        auto uni = [&]() { return (float)((next() >> 40) * (1.0 / 16777216.0)); };
        auto mag = [&]() { // a magnitude across many decades
            static const float dec[] = {1e-12f, 1e-6f, 1e-3f, 0.1f, 1.0f, 1.0f, 1.0f, 10.0f, 1e3f, 1e6f};
            return dec[next() % 10] * (0.1f + uni());
        };
        const float scale = mag();
        V3 lo, hi, o, d;
        float* L = &lo.x; float* Hh = &hi.x; float* O = &o.x; float* D = &d.x;
        for (int k = 0; k < 3; k++) {
            const float c = (uni() - 0.5f) * 4.0f * scale, e = (next() % 8 == 0) ? 0.0f : uni() * scale;
            L[k] = c - e;
            Hh[k] = c + e;
            const unsigned pick = (unsigned)(next() % 10);
            if (pick == 0) O[k] = L[k];                                  // on a face
            else if (pick == 1) O[k] = Hh[k];
            else if (pick == 2) O[k] = std::nextafter(L[k], -1e30f);     // one ulp outside / inside
            else if (pick == 3) O[k] = std::nextafter(Hh[k], 1e30f);
            else if (pick == 4) O[k] = c + (uni() - 0.5f) * 1e3f * scale; // far away
            else O[k] = c + (uni() - 0.5f) * 6.0f * scale;
            const unsigned dp = (unsigned)(next() % 12);
            float dv = (uni() - 0.5f) * 2.0f;
            if (dp == 0) dv = 0.0f;
            else if (dp == 1) dv *= 1e-20f;
            else if (dp == 2) dv *= 1e-35f;
            else if (dp == 3) dv *= 1e20f;
            else if (dp == 4) dv *= 1e-7f;
            D[k] = dv;
        }
        if (next() % 4 == 0) { // aim at the box, so that hits are common
            const V3 tgt = mk3(lo.x + (hi.x - lo.x) * uni(), lo.y + (hi.y - lo.y) * uni(), lo.z + (hi.z - lo.z) * uni());
            d = tgt - o;
        }
        const SlabFast F = slabFastBegin(d);
        float up = 0.0f;
        if (!slabFastHit(F, lo, hi, o, up)) continue;
        claims++;
        float te = 0.0f;
        if (!slabTest(lo, hi, o, d, __int_as_float(0x7f800000), te) || !(te <= up)) bad++;
        else if (up > te * 1.00001f + 1e-30f) loose++;
    }
    out[0] = n; out[1] = claims; out[2] = bad; out[3] = loose;
    return 0;
}

} // extern "C"
