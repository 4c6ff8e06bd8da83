// Host BVH builder: the reference's split rule applied on index ranges (no per-node mesh copies).
// Replaces BoundingVolumeHierarchy::BoundingVolumeHierarchy(Scene*) + getSubNodes/createTree/getBoundingBoxFromMeshes
// (src/bounding_volume_hierarchy.cpp:42-76, 88-207, 235-389). Emits BFS-numbered nodes (children adjacent) and the
// triangles of every leaf in the leaf's own visiting order, ready to be flattened into 32-byte device nodes.
#pragma once
#include <cstdint>
#include <vector>

namespace cgrt {

struct MeshView {
    const float* vertices;     // [nv][6]  p.xyz n.xyz
    const uint32_t* triangles; // [nt][3]  mesh-local vertex indices
    int32_t nv, nt;
    int32_t triOffset; // global id of this mesh's first triangle
};

struct HostNode {
    float lo[3], hi[3];
    int32_t child0, child1; // -1 for leaves; child1 == child0 + 1 always (createTree pushes the pair consecutively)
    int32_t firstTri, triCount; // range in BuiltBVH::leafTris (leaves only)
    int32_t level;
    int32_t isLeaf;
};

struct LeafTri {
    int32_t mesh;
    int32_t tri; // mesh-local triangle index
};

// 8-wide node of the culling sub-trees that refine the (large) leaves of the reference tree. Boxes are pre-expanded so that
// culling is conservative with respect to the reference's floating-point accept test (see buildLeafSubTrees). Unused child
// slots have an inverted box (never hit). Child ids are already in the traversal's encoding (cgrt_device.cuh):
// CGRT_SUB | wide node index, or CGRT_SUB | CGRT_TRI | (count-1) << 26 | first position.
struct WideNode {
    float lo[8][3], hi[8][3];
    uint32_t id[8];
};

struct BuiltBVH {
    std::vector<HostNode> nodes;
    std::vector<LeafTri> leafTris;   // after buildLeafSubTrees: permuted inside each reference leaf (sub-tree order)
    std::vector<int32_t> leafRank;   // per position: rank of that triangle in the reference's own leaf order
    std::vector<LeafTri> leafTrisReferenceOrder; // the reference's visiting order (intersectLeaf), kept for introspection
    std::vector<WideNode> wide;
    std::vector<int32_t> wideRoot;   // per reference node: root of its sub-tree in `wide`, -1 = scan the leaf
    // speculative ("fast") traversal, buildFastTree: one 8-wide conservative tree over ALL triangles (top = the reference tree
    // collapsed three levels at a time, bottom = the leaf sub-trees above) + what certification of its result needs
    uint32_t fastRoot = 0;             // id of the root in the traversal's encoding, 0 = no fast tree
    std::vector<int32_t> parent;       // per reference node: parent index (-1 for the root)
    std::vector<int32_t> triLeafNode;  // per position: reference leaf that holds the triangle
    std::vector<int32_t> fastOrder;    // fast-tree triangle order: fast position -> position (the fast tree's leaves are ranges of
                                       // THIS order; identity when the fast tree reuses the reference leaves' sub-trees)
    std::vector<int32_t> alwaysTest;   // positions of triangles whose accept region cannot be bounded (extreme slivers, non-finite):
                                       // not covered by the fast tree's boxes, tested for every ray that enters the tree
    int numLevels = 0;
};

// maxDepth: the reference literal is 12 (bvh.cpp:48); leaves are nodes at level maxDepth-1 or single-mesh/single-triangle nodes.
void buildReferenceBVH(const std::vector<MeshView>& meshes, int maxDepth, BuiltBVH& out);

// Refine every reference leaf with at least `minLeafForSubTree` triangles by an 8-wide culling tree (median splits of the
// centroids along the longest axis, three binary levels collapsed into one node, at most `subLeafSize` <= 8 triangles per
// sub-leaf). The reference tree itself is untouched: the traversal still visits reference nodes in the reference's order
// with the reference's exact box decisions; inside a reference leaf the sub-tree only decides which triangles need the exact
// test. Triangles whose accept region cannot be bounded tightly (non-finite coordinates, minimum angle below ~0.01 rad) get
// an unbounded box, i.e. they are always tested.
void buildLeafSubTrees(const std::vector<MeshView>& meshes, BuiltBVH& bvh, int minLeafForSubTree = 8, int subLeafSize = 6);

// The speculative traversal's tree (after buildLeafSubTrees): appends the collapsed top levels to bvh.wide and fills
// fastRoot / parent / triLeafNode. Boxes are unions of the triangles' conservative boxes, so the tolerant slab test can never
// cull a triangle the reference could accept, wherever it sits in the reference tree.
// sah = false: the reference tree collapsed three levels at a time on top of the leaf sub-trees (shares their triangle order).
// sah = true : an independent binned-SAH tree over all triangles with its own triangle order (fastOrder); the certificate
//              still walks the REFERENCE tree (parent / triLeafNode), so results do not depend on this choice.
void buildFastTree(const std::vector<MeshView>& meshes, BuiltBVH& bvh, bool sah = false);

} // namespace cgrt
