// Device-side BVH traversal shared by every kernel: the reference's exact visiting order over the reference tree, with the
// (large) reference leaves refined by conservative culling sub-trees. Results are identical to the reference's sequential
// leaf scan by construction; see the comments on leaf evaluation below.
#pragma once
#include "rt_math.cuh"
#include "cgrt_kernels.h"

RT_DEV int f2i(float f) { return __float_as_int(f); }
RT_DEV float i2f(int i) { return __int_as_float(i); }

struct TraceResult {
    float t;      // ray.t after the query
    int tri;      // position (leaf-ordered arrays) of the last accepted triangle, -1 none
    int sphere;   // index of the last accepted sphere (closer than every triangle), -1 none
    V3 sphereN;   // its normal
};

// ---- leaf evaluation ----------------------------------------------------------------------------------------------------------
// intersectLeaf (src/bounding_volume_hierarchy.cpp:535-553) scans the leaf's triangles in order and accepts a triangle iff
// its plane distance tt is >= 0 and < the CURRENT ray.t and the hit point is inside (src/ray_tracing.cpp:40-114). The state
// after the scan is therefore: among the triangles that are acceptable against the ray.t at leaf ENTRY, the one with the
// smallest tt, ties broken by the earliest position in the leaf ("rank"). That characterisation does not depend on the order
// in which the triangles are examined, which is what lets the sub-tree skip triangles that cannot be accepted.
// One quirk: when the origin lies exactly in the triangle's plane (dot(o,n) == D) the reference takes t = 0 WITHOUT comparing
// against ray.t (:43-47), so among such "shortcut" candidates the LAST one in leaf order wins, and it beats every other.
struct LeafBest {
    float t;      // best distance so far (initially ray.t at leaf entry)
    int pos;      // position of the best triangle, -1 = none accepted in this leaf yet
    int rank;     // its rank in the reference's leaf order
    bool shortcut;
};

// Exact accept arithmetic for one triangle (same expression trees as planeTest + pointInTriangleDev), folded into `best`.
// Returns true iff the triangle became the new best.
RT_DEV bool leafCandidate(const DevScene& S, int i, const V3& o, const V3& d, LeafBest& best)
{
    const float4 pl = __ldg(S.triPl + i);
    const V3 n = mk3(pl);
    const float on = dot3(o, n);
    float tt;
    bool shortcut = false;
    if (on == pl.w) {
        tt = 0.0f;
        shortcut = true;
    } else {
        const float denominator = dot3(d, n);
        if (denominator == 0) return false;
        tt = (pl.w - on) / denominator;
        if (tt < 0) return false;
        // cannot win: farther than the best, or a shortcut candidate already holds the leaf, or equal to ray.t at leaf
        // entry (rejected by the reference's `t >= ray.t`)
        if (!(tt <= best.t)) return false;
        if (best.shortcut) return false;
        if (tt == best.t && best.pos < 0) return false;
    }
    const float4 v0 = __ldg(S.triV0 + i), v1 = __ldg(S.triV1 + i), v2 = __ldg(S.triV2 + i);
    const int rank = f2i(v2.w);
    if (shortcut) {
        if (best.shortcut && rank < best.rank) return false; // the last shortcut candidate in leaf order wins
    } else if (tt == best.t && rank > best.rank) {
        return false; // equally close: the reference keeps the one it reached first
    }
    const V3 p = o + d * tt;
    if (!pointInTriangleDev(mk3(v0), mk3(v1), mk3(v2), n, p)) return false;
    best.t = tt;
    best.pos = i;
    best.rank = rank;
    best.shortcut = shortcut;
    return true;
}

// Tolerant slab test against a pre-expanded sub-tree box with the per-ray reciprocal direction. Never reports a miss for a box
// that contains a point the reference could accept at a distance <= bt (bvh_build.cpp: soundness argument).
RT_DEV bool slabLoose(const float4& lo, const float4& hi, const V3& o, const V3& inv, float bt, float& tNear)
{
    const float t0x = (lo.x - o.x) * inv.x, t1x = (hi.x - o.x) * inv.x;
    const float t0y = (lo.y - o.y) * inv.y, t1y = (hi.y - o.y) * inv.y;
    const float t0z = (lo.z - o.z) * inv.z, t1z = (hi.z - o.z) * inv.z;
    const float tIn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    const float tOut = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    tNear = tIn;
    const float slack = 1.000001f;
    return !(tOut < 0.0f || tIn > tOut * slack || tIn > bt * slack);
}

#define CGRT_SUBSTACK 40
#define CGRT_SUBFLAG 0x40000000 // stack / current-node ids with this bit set refer to S.subNodes

// ---- the production traversal: one node per step, reference nodes and sub-tree nodes on ONE stack, resumable --------------
// Same visiting order, pruning and accept arithmetic as traverseStrict below (which documents the mapping to the reference's
// functions); the difference is purely structural. Every step handles exactly one node - a reference inner node (exact
// ordered logic), a reference leaf without sub-tree (sequential scan), a sub-tree inner node (tolerant slab tests) or a
// sub-tree leaf (exact triangle tests) - so that the lanes of a warp advance in lock-step instead of waiting for each other's
// nested loops, and so that a lane whose ray is finished can be handed a new ray between steps (persistent warps,
// cgrt_kernels.cu). Stack entries carry a key: for a reference sibling the entry distance tSecond (skipped iff
// ray.t < tSecond, intersectChildrenHierarchically), for a sub-tree node its tolerant entry distance (skipped iff it lies
// beyond the leaf's best distance). A reference leaf's result is committed when its sub-tree entries are gone.
struct Trav {
    V3 o, d, inv;
    float t;
    int hitTri;
    int sp, node, subBase;
    bool useSub, inLeaf;
    bool fastRef; // reference-node box tests may go through the filtered (division-free) path for this ray
    LeafBest best;
};
// the traversal stack lives outside the struct so that the scalars above stay in registers
struct TravStack {
    int n[CGRT_STACK + CGRT_SUBSTACK];
    float t[CGRT_STACK + CGRT_SUBSTACK];
};

enum { TRAV_CONTINUE = 0, TRAV_DONE = 1, TRAV_FIRED = 2 };

// intersectDataStructure (bvh.cpp:831-844): returns true iff the tree has to be traversed for this ray.
RT_DEV bool travBegin(const DevScene& S, Trav& T, const V3& o, const V3& d, float tIn)
{
    T.o = o;
    T.d = d;
    T.t = tIn;
    T.hitTri = -1;
    T.sp = 0;
    T.node = 0;
    T.subBase = 0;
    T.inLeaf = false;
    T.best.t = tIn; T.best.pos = -1; T.best.rank = -1; T.best.shortcut = false;
    if (S.nNodes <= 0) return false;
    const float4 rq0 = __ldg(S.nodes + 0), rq1 = __ldg(S.nodes + 1);
    bool enter = startsInBox(o, mk3(rq0), mk3(rq1));
    if (!enter) {
        float tmp;
        enter = slabTest(mk3(rq0), mk3(rq1), o, d, tIn, tmp);
    }
    if (!enter) return false;
    // per-ray data of the tolerant sub-tree slab test; rays with extreme direction components scan leaves instead
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const bool okx = (ax == 0.0f) || (ax >= 1e-20f && ax <= 1e20f);
    const bool oky = (ay == 0.0f) || (ay >= 1e-20f && ay <= 1e20f);
    const bool okz = (az == 0.0f) || (az >= 1e-20f && az <= 1e20f);
    const bool fin = fabsf(o.x) <= 1e30f && fabsf(o.y) <= 1e30f && fabsf(o.z) <= 1e30f; // false for NaN
    T.useSub = S.subNodes != nullptr && okx && oky && okz && fin;
    T.inv.x = ax == 0.0f ? 1e30f : 1.0f / d.x;
    T.inv.y = ay == 0.0f ? 1e30f : 1.0f / d.y;
    T.inv.z = az == 0.0f ? 1e30f : 1.0f / d.z;
    // filtered reference box tests need finite non-zero reciprocals and no overflow in (box - o) * inv
    T.fastRef = ax >= 1e-15f && ax <= 1e15f && ay >= 1e-15f && ay <= 1e15f && az >= 1e-15f && az <= 1e15f &&
                fabsf(o.x) <= 1e15f && fabsf(o.y) <= 1e15f && fabsf(o.z) <= 1e15f;
    return true;
}

// ---- filtered evaluation of the reference's exact box test ----------------------------------------------------------------
// slabTest (src/ray_tracing.cpp:162-200) needs six IEEE divisions; almost every call is decided with a wide margin, and its
// numeric result matters only when two children have to be ordered or a pending sibling is pruned. The filter evaluates the
// slab distances as (box - o) * fl(1/d): each such value differs from the reference's correctly rounded quotient by at most
// 3 ulp, so with rho = 1e-6 relative (+ tau absolute, for the subnormal range) slack every comparison of the reference is
// either decided with certainty or declared ambiguous; ambiguous cases fall back to slabTest itself. Outcomes are therefore
// exactly the reference's, the divisions are just not executed when they cannot matter.
// Measured on B200 (profiles/r01_tuning.md): the filter removes ~25 % of the executed instructions of a reference-node step
// but does not shorten the frame (the kernels are bound by divergence / instruction fetch, not by the division pipe), so it
// is compiled out by default; -DCGRT_USE_FILTER=1 enables it (the parity suite passes either way).
#ifndef CGRT_USE_FILTER
#define CGRT_USE_FILTER 0
#endif
#define CGRT_RHO 1e-6f
#define CGRT_TAU 1e-30f
RT_DEV float errBound(float x) { return fabsf(x) * CGRT_RHO + CGRT_TAU; }

// the exact test, kept out of line: it is the rare fallback of the filter and would otherwise be inlined six times
CGRT_RARE bool slabTestOutOfLine(const float4& lo, const float4& hi, const V3& o, const V3& d, float rayT, float& tHit)
{
    return slabTest(mk3(lo), mk3(hi), o, d, rayT, tHit);
}

// hit: the reference's boolean; t: its distance (exact when `exact`, otherwise within errBound(t) of it)
RT_DEV void boxFiltered(const float4& lo, const float4& hi, const V3& o, const V3& d, const V3& inv, bool fast, float rayT,
                        bool& hit, float& t, bool& exact)
{
    if (CGRT_USE_FILTER && fast) {
        const float q0x = (lo.x - o.x) * inv.x, q1x = (hi.x - o.x) * inv.x;
        const float q0y = (lo.y - o.y) * inv.y, q1y = (hi.y - o.y) * inv.y;
        const float q0z = (lo.z - o.z) * inv.z, q1z = (hi.z - o.z) * inv.z;
        const float tin = fmaxf(fmaxf(fminf(q0x, q1x), fminf(q0y, q1y)), fminf(q0z, q1z));
        const float tout = fminf(fminf(fmaxf(q0x, q1x), fmaxf(q0y, q1y)), fmaxf(q0z, q1z));
        const float ein = errBound(tin), eout = errBound(tout);
        if (tout + eout < 0.0f || tin - ein > tout + eout) { // certainly `tOut < 0` or `tIn > tOut`
            hit = false;
            return;
        }
        float cur, ecur;
        bool branch = false;
        if (tin + ein < 0.0f) { cur = tout; ecur = eout; branch = true; }        // certainly tIn < 0: inside the slabs
        else if (tin - ein >= 0.0f) { cur = tin; ecur = ein; branch = true; }   // certainly tIn >= 0
        if (branch) {
            if (cur - ecur >= rayT) { // certainly `currentT >= ray.t`
                hit = false;
                return;
            }
            if (tin + ein <= tout - eout && tout - eout >= 0.0f && cur + ecur < rayT) { // certainly a hit
                hit = true;
                t = cur;
                exact = false;
                return;
            }
        }
    }
    float te = 0.0f;
    hit = slabTestOutOfLine(lo, hi, o, d, rayT, te);
    t = te;
    exact = true;
}

// the reference's distance of a box that is known to be hit (used when an approximate distance is not decisive)
CGRT_RARE float boxExactT(const float4& lo, const float4& hi, const V3& o, const V3& d)
{
    float te = 0.0f;
    slabTest(mk3(lo), mk3(hi), o, d, __int_as_float(0x7f800000), te);
    return te;
}

// Node classes of the state machine. The current node id encodes its class: reference nodes are plain indices into
// S.nodes; CGRT_SUBFLAG marks a sub-tree inner node (index into S.subNodes); CGRT_SUBFLAG|CGRT_LEAFFLAG marks a sub-tree
// leaf and carries its triangle range directly (bit 28 = count-1, bits 0..27 = first position), so a sub-tree leaf costs
// no node fetch.
#define CGRT_LEAFFLAG 0x20000000
#define CGRT_KEYAPPROX 0x10000000 // reference stack entry whose key is a filtered (approximate) box distance
#define CGRT_LEAFCNT_SHIFT 28
#define CGRT_POS_MASK 0x0fffffff
enum { CLS_REF = 0, CLS_SUBINNER = 1, CLS_SUBLEAF = 2, CLS_NONE = 3 };
RT_DEV int travClass(int node)
{
    return !(node & CGRT_SUBFLAG) ? CLS_REF : ((node & CGRT_LEAFFLAG) ? CLS_SUBLEAF : CLS_SUBINNER);
}
RT_DEV int subChildId(int index, const float4& lo, const float4& hi)
{
    const int b = f2i(hi.w);
    if (b == 0) return index | CGRT_SUBFLAG;
    return CGRT_SUBFLAG | CGRT_LEAFFLAG | ((b - 1) << CGRT_LEAFCNT_SHIFT) | f2i(lo.w);
}

// pop: next pending node that is not pruned, committing the reference leaf when its sub-tree entries are gone
RT_DEV int travPop(const DevScene& S, Trav& T, TravStack& K)
{
    const float slack = 1.000001f;
    while (true) {
        if (T.inLeaf && T.sp == T.subBase) {
            if (T.best.pos >= 0) { T.t = T.best.t; T.hitTri = T.best.pos; }
            T.inLeaf = false;
        }
        if (T.sp == 0) return TRAV_DONE;
        T.sp--;
        const int n = K.n[T.sp];
        const float key = K.t[T.sp];
        if (n & CGRT_SUBFLAG) {
            if (key > T.best.t * slack) continue;
            T.node = n;
            return TRAV_CONTINUE;
        }
        // reference sibling: skipped iff ray.t < tSecond (intersectChildrenHierarchically, bvh.cpp:572-595)
        const int ni = n & ~CGRT_KEYAPPROX;
        if (n & CGRT_KEYAPPROX) {
            const float e = errBound(key);
            if (T.t < key - e) continue;
            if (!(T.t >= key + e)) { // not decisive: compare with the reference's exact distance of that box
                const float tS = boxExactT(__ldg(S.nodes + 2 * ni), __ldg(S.nodes + 2 * ni + 1), T.o, T.d);
                if (T.t < tS) continue;
            }
        } else {
            if (T.t < key) continue;
        }
        T.node = ni;
        return TRAV_CONTINUE;
    }
}

// one step on a reference node
template <bool ANY>
RT_DEV int travStepRef(const DevScene& S, Trav& T, TravStack& K, float eps, float maxDist)
{
    const V3 o = T.o, d = T.d;
    const int node = T.node;
    const float4 q0 = __ldg(S.nodes + 2 * node), q1 = __ldg(S.nodes + 2 * node + 1);
    const uint32_t a = (uint32_t)f2i(q0.w), b = (uint32_t)f2i(q1.w);
    if (b != 0u) {
        // ---- reference leaf
        T.best.t = T.t; T.best.pos = -1; T.best.rank = -1; T.best.shortcut = false;
        const int sr = T.useSub ? __ldg(S.subRoot + node) : -1;
        if (sr >= 0) {
            T.inLeaf = true;
            T.subBase = T.sp;
            T.node = sr | CGRT_SUBFLAG; // the root of a sub-tree is always an inner node (leaves below 8 triangles have none)
            return TRAV_CONTINUE;
        }
        const uint32_t end = a + b;
#pragma unroll 1
        for (uint32_t i = a; i < end; i++) {
            if (leafCandidate(S, (int)i, o, d, T.best)) {
                if (ANY && !(T.best.t + eps >= maxDist)) return TRAV_FIRED;
            }
        }
        if (T.best.pos >= 0) { T.t = T.best.t; T.hitTri = T.best.pos; }
        return travPop(S, T, K);
    }
    // ---- reference inner node: intersectNonLeaf + intersectDeeper, exact arithmetic
    const int L = (int)a, Rn = (int)a + 1;
    const float4 l0 = __ldg(S.nodes + 2 * L), l1 = __ldg(S.nodes + 2 * L + 1);
    const float4 r0 = __ldg(S.nodes + 2 * L + 2), r1 = __ldg(S.nodes + 2 * L + 3);
    bool hL, hR, exL = true, exR = true;
    float tL = -1.0f, tR = -1.0f;
    boxFiltered(l0, l1, o, d, T.inv, T.fastRef, T.t, hL, tL, exL);
    boxFiltered(r0, r1, o, d, T.inv, T.fastRef, T.t, hR, tR, exR);
    const bool inL = startsInBox(o, mk3(l0), mk3(l1));
    const bool inR = startsInBox(o, mk3(r0), mk3(r1));
    int first = -1, second = -1;
    float tS = -1.0f;
    bool keyApprox = false;
    if (inL && inR) { // both visited unconditionally, left operand of `|` first (g++ order)
        first = L; second = Rn; tS = -1.0f;
    } else if (inL) {
        first = L;
        if (hR) { second = Rn; tS = tR; keyApprox = !exR; }
    } else if (inR) {
        first = Rn;
        if (hL) { second = L; tS = tL; keyApprox = !exL; }
    } else if (hL && hR) {
        // nearer child first: `tLeft < tRight` (bvh.cpp:626); decide from the approximate distances when they are
        // separated by more than their error bounds, otherwise from the exact ones
        bool leftFirst;
        if (tL + errBound(tL) < tR - errBound(tR) && !(exL && exR)) leftFirst = true;
        else if (tL - errBound(tL) >= tR + errBound(tR) && !(exL && exR)) leftFirst = false;
        else {
            if (!exL) { tL = boxExactT(l0, l1, o, d); exL = true; }
            if (!exR) { tR = boxExactT(r0, r1, o, d); exR = true; }
            leftFirst = tL < tR;
        }
        if (leftFirst) { first = L; second = Rn; tS = tR; keyApprox = !exR; }
        else { first = Rn; second = L; tS = tL; keyApprox = !exL; }
    } else if (hL) {
        first = L;
    } else if (hR) {
        first = Rn;
    }
    if (second >= 0) {
        K.n[T.sp] = keyApprox ? (second | CGRT_KEYAPPROX) : second;
        K.t[T.sp] = tS;
        T.sp++;
    }
    if (first >= 0) {
        T.node = first;
        return TRAV_CONTINUE;
    }
    return travPop(S, T, K);
}

// one step on a sub-tree inner node: tolerant slab tests, nearer child first
RT_DEV int travStepSubInner(const DevScene& S, Trav& T, TravStack& K)
{
    const int sn = T.node & CGRT_POS_MASK;
    const int a = f2i(__ldg(S.subNodes + 2 * sn).w);
    const float4 l0 = __ldg(S.subNodes + 2 * a), l1 = __ldg(S.subNodes + 2 * a + 1);
    const float4 r0 = __ldg(S.subNodes + 2 * a + 2), r1 = __ldg(S.subNodes + 2 * a + 3);
    float tL, tR;
    const bool hL = slabLoose(l0, l1, T.o, T.inv, T.best.t, tL);
    const bool hR = slabLoose(r0, r1, T.o, T.inv, T.best.t, tR);
    const int idL = subChildId(a, l0, l1), idR = subChildId(a + 1, r0, r1);
    if (hL && hR) {
        const bool leftFirst = tL <= tR;
        K.n[T.sp] = leftFirst ? idR : idL;
        K.t[T.sp] = leftFirst ? tR : tL;
        T.sp++;
        T.node = leftFirst ? idL : idR;
        return TRAV_CONTINUE;
    }
    if (hL || hR) {
        T.node = hL ? idL : idR;
        return TRAV_CONTINUE;
    }
    return travPop(S, T, K);
}

// one step on a sub-tree leaf: exact tests of the few triangles that survived the culling
template <bool ANY>
RT_DEV int travStepSubLeaf(const DevScene& S, Trav& T, TravStack& K, float eps, float maxDist)
{
    const int first = T.node & CGRT_POS_MASK;
    const int count = ((T.node >> CGRT_LEAFCNT_SHIFT) & 1) + 1;
#pragma unroll 1
    for (int i = first; i < first + count; i++) {
        if (leafCandidate(S, i, T.o, T.d, T.best)) {
            if (ANY && !(T.best.t + eps >= maxDist)) return TRAV_FIRED;
        }
    }
    return travPop(S, T, K);
}

template <bool ANY>
RT_DEV int travStep(const DevScene& S, Trav& T, TravStack& K, float eps, float maxDist)
{
    const int cls = travClass(T.node);
    if (cls == CLS_REF) return travStepRef<ANY>(S, T, K, eps, maxDist);
    if (cls == CLS_SUBINNER) return travStepSubInner(S, T, K);
    return travStepSubLeaf<ANY>(S, T, K, eps, maxDist);
}

// After the tree: the sphere loop of BoundingVolumeHierarchy::intersect (bvh.cpp:878-879) and the result record.
// `state` is TRAV_FIRED when the any-hit predicate already fired inside the tree. Returns hit (closest) / shadowed (ANY).
template <bool ANY>
RT_DEV bool travFinish(const DevScene& S, Trav& T, int state, float eps, float maxDist, TraceResult& R)
{
    R.sphere = -1;
    if (ANY && state == TRAV_FIRED) {
        R.t = T.best.t;
        R.tri = T.best.pos;
        return true;
    }
    float t = T.t;
    for (int s = 0; s < S.nSpheres; s++) {
        const float4 c = __ldg(S.spheres + 3 * s);
        float ts;
        V3 nn;
        if (sphereTest(mk3(c), c.w, T.o, T.d, t, ts, nn)) {
            t = ts;
            R.sphere = s;
            R.sphereN = nn;
            if (ANY && !(ts + eps >= maxDist)) {
                R.t = t;
                R.tri = T.hitTri;
                return true;
            }
        }
    }
    R.t = t;
    R.tri = T.hitTri;
    if (ANY) return false;
    return T.hitTri >= 0 || R.sphere >= 0;
}

// one ray, start to finish (batch entry points that do not use persistent warps)
template <bool ANY>
RT_DEV bool traverseFast(const DevScene& S, const V3& o, const V3& d, float tIn, float eps, float maxDist, TraceResult& R)
{
    Trav T;
    TravStack K;
    int state = TRAV_DONE;
    if (travBegin(S, T, o, d, tIn)) {
        do {
            state = travStep<ANY>(S, T, K, eps, maxDist);
        } while (state == TRAV_CONTINUE);
    }
    return travFinish<ANY>(S, T, state, eps, maxDist, R);
}

// Closest-hit traversal in the reference's exact visiting order:
//   intersect            src/bounding_volume_hierarchy.cpp:850-881
//   intersectDataStructure :831-844   enter iff origin strictly inside root box OR slab test passes against ray.t
//   intersectNonLeaf     :715-736     slab-test BOTH children against the current ray.t (tLeft/tRight, -1 = miss)
//   intersectDeeper      :679-701     classify by startsInBox
//   intersectChildrenHierarchically :572-595, intersectRayThatStartsOutsideBoxes :611-635 (with the intended `return`s)
//   intersectLeaf        :535-553     (through leafCandidate, see above)
// The recursion is unrolled onto an explicit stack of (node, tSecond): the pending sibling is skipped at pop time iff
// ray.t < tSecond, which equals the reference's `hitFirst && ray.t < tSecond` because tSecond < ray.t held when it was pushed.
// ANY = true adds an early exit as soon as an accepted hit satisfies the shadow predicate !(t + eps >= maxDist)
// (pointInShadow, src/main.cpp:104-135); later accepted hits can only be closer, so the answer equals the closest-hit one.
// This variant scans every leaf sequentially (no sub-tree) and, with COUNT = true, counts the reference's box / triangle tests;
// the kernels use it for the counting passes, traverseFast above for production.
template <bool ANY, bool COUNT>
RT_DEV bool traverseStrict(const DevScene& S, const V3& o, const V3& d, float tIn, float eps, float maxDist, TraceResult& R,
                           uint32_t& nBox, uint32_t& nTri)
{
    float t = tIn;
    int hitTri = -1;
    R.sphere = -1;
    if (S.nNodes > 0) {
        float4 q0 = __ldg(S.nodes + 0), q1 = __ldg(S.nodes + 1);
        bool enter = startsInBox(o, mk3(q0), mk3(q1));
        if (!enter) {
            float tmp;
            if (COUNT) nBox++;
            enter = slabTest(mk3(q0), mk3(q1), o, d, t, tmp);
        }
        if (enter) {
            int stN[CGRT_STACK];
            float stT[CGRT_STACK];
            int sp = 0;
            int cur = 0;
            uint32_t a = (uint32_t)f2i(q0.w), b = (uint32_t)f2i(q1.w);
            while (true) {
                if (b != 0u) {
                    // ---- intersectLeaf
                    LeafBest best;
                    best.t = t;
                    best.pos = -1;
                    best.rank = -1;
                    best.shortcut = false;
                    if (COUNT) nTri += b;
                    {
                        const uint32_t end = a + b;
                        for (uint32_t i = a; i < end; i++) {
                            if (leafCandidate(S, (int)i, o, d, best)) {
                                if (ANY && !(best.t + eps >= maxDist)) {
                                    R.t = best.t;
                                    R.tri = best.pos;
                                    return true;
                                }
                            }
                        }
                    }
                    if (best.pos >= 0) {
                        t = best.t;
                        hitTri = best.pos;
                    }
                    // ---- return to the nearest pending sibling that is not pruned
                    bool found = false;
                    while (sp > 0) {
                        sp--;
                        if (t < stT[sp]) continue;
                        cur = stN[sp];
                        q0 = __ldg(S.nodes + 2 * cur);
                        q1 = __ldg(S.nodes + 2 * cur + 1);
                        a = (uint32_t)f2i(q0.w);
                        b = (uint32_t)f2i(q1.w);
                        found = true;
                        break;
                    }
                    if (!found) break;
                } else {
                    // ---- intersectNonLeaf + intersectDeeper
                    const int L = (int)a, Rn = (int)a + 1;
                    const float4 l0 = __ldg(S.nodes + 2 * L), l1 = __ldg(S.nodes + 2 * L + 1);
                    const float4 r0 = __ldg(S.nodes + 2 * L + 2), r1 = __ldg(S.nodes + 2 * L + 3);
                    float tL = -1.0f, tR = -1.0f, tmp;
                    if (COUNT) nBox += 2;
                    if (slabTest(mk3(l0), mk3(l1), o, d, t, tmp)) tL = tmp;
                    if (slabTest(mk3(r0), mk3(r1), o, d, t, tmp)) tR = tmp;
                    const bool inL = startsInBox(o, mk3(l0), mk3(l1));
                    const bool inR = startsInBox(o, mk3(r0), mk3(r1));
                    int first = -1, second = -1;
                    float tS = -1.0f;
                    if (inL && inR) { // both visited unconditionally, left operand of `|` first (g++ order)
                        first = L; second = Rn; tS = -1.0f;
                    } else if (inL) {
                        first = L;
                        if (!(tR < 0)) { second = Rn; tS = tR; }
                    } else if (inR) {
                        first = Rn;
                        if (!(tL < 0)) { second = L; tS = tL; }
                    } else {
                        if (tL < 0 && tR < 0) {
                        } else if (tL < 0) {
                            first = Rn;
                        } else if (tR < 0) {
                            first = L;
                        } else if (tL < tR) {
                            first = L; second = Rn; tS = tR;
                        } else {
                            first = Rn; second = L; tS = tL;
                        }
                    }
                    if (second >= 0) {
                        stN[sp] = second;
                        stT[sp] = tS;
                        sp++;
                    }
                    if (first >= 0) {
                        cur = first;
                        if (first == L) { a = (uint32_t)f2i(l0.w); b = (uint32_t)f2i(l1.w); }
                        else { a = (uint32_t)f2i(r0.w); b = (uint32_t)f2i(r1.w); }
                    } else {
                        bool found = false;
                        while (sp > 0) {
                            sp--;
                            if (t < stT[sp]) continue;
                            cur = stN[sp];
                            q0 = __ldg(S.nodes + 2 * cur);
                            q1 = __ldg(S.nodes + 2 * cur + 1);
                            a = (uint32_t)f2i(q0.w);
                            b = (uint32_t)f2i(q1.w);
                            found = true;
                            break;
                        }
                        if (!found) break;
                    }
                }
            }
        }
    }
    // ---- sphere loop, src/bounding_volume_hierarchy.cpp:878-879
    for (int s = 0; s < S.nSpheres; s++) {
        const float4 c = __ldg(S.spheres + 3 * s);
        float ts;
        V3 nn;
        if (sphereTest(mk3(c), c.w, o, d, t, ts, nn)) {
            t = ts;
            R.sphere = s;
            R.sphereN = nn;
            if (ANY && !(ts + eps >= maxDist)) {
                R.t = t;
                R.tri = hitTri;
                return true;
            }
        }
    }
    R.t = t;
    R.tri = hitTri;
    if (ANY) return false;
    return hitTri >= 0 || R.sphere >= 0;
}
