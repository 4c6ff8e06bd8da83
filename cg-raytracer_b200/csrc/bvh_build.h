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

struct BuiltBVH {
    std::vector<HostNode> nodes;
    std::vector<LeafTri> leafTris;
    int numLevels = 0;
};

// maxDepth: the reference literal is 12 (bvh.cpp:48); leaves are nodes at level maxDepth-1 or single-mesh/single-triangle nodes.
void buildReferenceBVH(const std::vector<MeshView>& meshes, int maxDepth, BuiltBVH& out);

} // namespace cgrt
