/* cgrt_host_c.h — C access to the host-side data formats either side of the hot path:
 * the OBJ/MTL loader + scene presets that feed it and the BMP writer that consumes the framebuffer. Pure host code
 * (no CUDA): these produce / consume the flat arrays of cgrt_scene_desc, they never compute intersections. */
#ifndef CGRT_HOST_C_H
#define CGRT_HOST_C_H
#include "cgrt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgrt_host_scene cgrt_host_scene; /* a loaded Scene (src/scene.h:53-60) + its flattened arrays */

/* device used by the free functions of the C++ mirror (intersectRayWith*, BoundingVolumeHierarchy(Scene*)) */
void cgrt_host_set_default_device(int device);

/* loadScene(SceneType, dataDir), src/scene.cpp:4-69. preset = "SingleTriangle" | "Cube" | "CornellBox" |
 * "CornellBoxSphericalLight" | "Monkey" | "Dragon" | "Spheres" | "Custom". "Dragon" falls back to the named procedural
 * stand-in when <data_dir>/dragon.obj is absent (it is not part of the reference checkout). */
int cgrt_host_scene_load_preset(const char* preset, const char* data_dir, cgrt_host_scene** out);
/* loadMesh(file, normalize), src/mesh.cpp:58-141 (+ centerAndScaleToUnitMesh :143-166) */
int cgrt_host_scene_load_obj(const char* path, int normalize, cgrt_host_scene** out);
/* the dragon stand-in at a chosen tessellation: 2*segments_u*segments_v triangles (default 340 x 128 = 87 040) */
int cgrt_host_scene_dragon_standin(int segments_u, int segments_v, cgrt_host_scene** out);
void cgrt_host_scene_destroy(cgrt_host_scene* hs);
/* pointers in *out stay valid until cgrt_host_scene_destroy */
int cgrt_host_scene_desc(const cgrt_host_scene* hs, cgrt_scene_desc* out);
/* returns the mesh count */
int64_t cgrt_host_scene_counts(const cgrt_host_scene* hs, int64_t* n_vertices, int64_t* n_triangles);
/* returns the number of point lights of the preset; fills up to cap */
int cgrt_host_scene_lights(const cgrt_host_scene* hs, cgrt_point_light* out, int cap);

/* Screen::writeBitmapToFile, src/screen.cpp:38-49: rgb = [H][W][3] float in Screen layout (row 0 = top) */
int cgrt_write_bmp(const char* path, const float* rgb, int width, int height);

#ifdef __cplusplus
}
#endif
#endif
