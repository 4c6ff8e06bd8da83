// Device-side scene layout and the strict (reference-order) BVH traversal shared by every kernel.
#pragma once
#include "rt_math.cuh"
#include "cgrt_kernels.h"


RT_DEV int f2i(float f) { return __float_as_int(f); }
RT_DEV float i2f(int i) { return __int_as_float(i); }

struct TraceResult {
    float t;      // ray.t after the query
    int tri;      // leaf-order index of the last accepted triangle, -1 none
    int sphere;   // index of the last accepted sphere (closer than every triangle), -1 none
    V3 sphereN;   // its normal
};

// Closest-hit traversal in the reference's exact visiting order:
//   intersect            src/bounding_volume_hierarchy.cpp:850-881
//   intersectDataStructure :831-844   enter iff origin strictly inside root box OR slab test passes against ray.t
//   intersectNonLeaf     :715-736     slab-test BOTH children against the current ray.t (tLeft/tRight, -1 = miss)
//   intersectDeeper      :679-701     classify by startsInBox
//   intersectChildrenHierarchically :572-595, intersectRayThatStartsOutsideBoxes :611-635 (with the intended `return`s)
//   intersectLeaf        :535-553
// The recursion is unrolled onto an explicit stack of (node, tSecond): the pending sibling is skipped at pop time iff
// ray.t < tSecond, which equals the reference's `hitFirst && ray.t < tSecond` because tSecond < ray.t held when it was pushed.
// ANY = true adds an early exit as soon as an accepted hit satisfies the shadow predicate !(t + eps >= maxDist)
// (pointInShadow, src/main.cpp:104-135); later accepted hits can only be closer, so the answer equals the closest-hit one.
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
            uint32_t a = (uint32_t)f2i(q0.w), b = (uint32_t)f2i(q1.w);
            while (true) {
                if (b != 0u) {
                    // ---- intersectLeaf: every triangle in leaf order, accept iff plane hit closer than ray.t and inside
                    const uint32_t end = a + b;
                    for (uint32_t i = a; i < end; i++) {
                        if (COUNT) nTri++;
                        const float4 pl = __ldg(S.triPl + i);
                        float tt;
                        if (planeTest(mk3(pl), pl.w, o, d, t, tt)) {
                            const float4 v0 = __ldg(S.triV0 + i), v1 = __ldg(S.triV1 + i), v2 = __ldg(S.triV2 + i);
                            const V3 p = o + d * tt;
                            if (pointInTriangleDev(mk3(v0), mk3(v1), mk3(v2), mk3(pl), p)) {
                                t = tt;
                                hitTri = (int)i;
                                if (ANY && !(tt + eps >= maxDist)) {
                                    R.t = t;
                                    R.tri = hitTri;
                                    return true;
                                }
                            }
                        }
                    }
                    // ---- return to the nearest pending sibling that is not pruned
                    bool found = false;
                    while (sp > 0) {
                        sp--;
                        if (t < stT[sp]) continue;
                        const int n = stN[sp];
                        q0 = __ldg(S.nodes + 2 * n);
                        q1 = __ldg(S.nodes + 2 * n + 1);
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
                        if (first == L) { a = (uint32_t)f2i(l0.w); b = (uint32_t)f2i(l1.w); }
                        else { a = (uint32_t)f2i(r0.w); b = (uint32_t)f2i(r1.w); }
                    } else {
                        bool found = false;
                        while (sp > 0) {
                            sp--;
                            if (t < stT[sp]) continue;
                            const int n = stN[sp];
                            q0 = __ldg(S.nodes + 2 * n);
                            q1 = __ldg(S.nodes + 2 * n + 1);
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
