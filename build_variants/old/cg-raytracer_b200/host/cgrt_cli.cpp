// Headless harness that replaces the reference's interactive main() (src/main.cpp:722-939, GUI out of scope): it drives the
// hot path exactly the way main() does — loadScene, BoundingVolumeHierarchy{&scene}, camera preset, renderRayTracing,
// Screen::writeBitmapToFile — through the reference-named C++ interface in cgrt_host.h.
//
//   cgrt_cli <data_dir> <preset> <width> <height> <trace_limit> <out.bmp> [x y]
//
// With x y it also casts the debug ray of the "R" key (main.cpp:747-753, 896-903) through bvh.intersect and prints the hit.
#include "cgrt_host.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

int main(int argc, char** argv)
{
    if (argc < 7) {
        std::fprintf(stderr, "usage: %s <data_dir> <preset> <width> <height> <trace_limit> <out.bmp> [x y]\n", argv[0]);
        return 2;
    }
    const std::string dataDir = argv[1], preset = argv[2];
    const int W = std::atoi(argv[3]), H = std::atoi(argv[4]), limit = std::atoi(argv[5]);
    static const struct { const char* name; SceneType type; } table[] = {
        {"SingleTriangle", SingleTriangle}, {"Cube", Cube}, {"CornellBox", CornellBox}, {"Monkey", Monkey},
        {"Dragon", Dragon}, {"Spheres", Spheres}, {"Custom", Custom}};
    SceneType type = SingleTriangle;
    bool found = false;
    for (const auto& e : table)
        if (preset == e.name) { type = e.type; found = true; }
    if (!found) {
        std::fprintf(stderr, "unknown preset %s\n", preset.c_str());
        return 2;
    }
    try {
        Scene scene = loadScene(type, dataDir);                                  // main.cpp:735
        BoundingVolumeHierarchy bvh{&scene};                                     // main.cpp:736
        Trackball camera{float(W) / float(H), glm::radians(50.0f), 3.0f};        // main.cpp:730
        camera.setCamera(glm::vec3(0.0f, 0.0f, 0.0f), glm::vec3(glm::radians(20.0f), glm::radians(20.0f), 0.0f), 3.0f); // :731
        Screen screen{glm::ivec2(W, H)};
        RenderOptions opt;
        opt.traceLimit = limit;
        const auto t0 = std::chrono::high_resolution_clock::now();               // main.cpp:792-796
        const RenderReport rep = renderRayTracing(scene, camera, bvh, screen, opt);
        const auto t1 = std::chrono::high_resolution_clock::now();
        std::cout << "Time to render image: " << std::chrono::duration<double, std::milli>(t1 - t0).count()
                  << " milliseconds" << std::endl;
        std::printf("levels=%d primary=%llu primary_hit=%llu shadow=%llu bounce=%llu launches=%llu device_ms=%.3f\n",
                    bvh.numLevels(), (unsigned long long)rep.primary, (unsigned long long)rep.primaryHit,
                    (unsigned long long)rep.shadow, (unsigned long long)rep.bounce, (unsigned long long)rep.kernelLaunches,
                    rep.deviceMs);
        screen.writeBitmapToFile(argv[6]);                                       // main.cpp:798
        if (argc >= 9) {
            const int x = std::atoi(argv[7]), y = std::atoi(argv[8]);
            Ray ray = camera.generateRay(glm::vec2(float(x) / W * 2.0f - 1.0f, float(y) / H * 2.0f - 1.0f));
            HitInfo hit;
            hit.normal = glm::vec3(0.0f);
            const bool h = bvh.intersect(ray, hit);
            std::printf("debug ray (%d,%d): hit=%d t=%.9g normal=(%.9g %.9g %.9g) kd=(%.9g %.9g %.9g)\n", x, y, int(h), ray.t,
                        hit.normal.x, hit.normal.y, hit.normal.z, hit.material.kd.x, hit.material.kd.y, hit.material.kd.z);
            std::printf("debug boxes at level 1: %zu\n", bvh.debugNodes(1).size());
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "cgrt_cli: %s\n", e.what());
        return 1;
    }
    return 0;
}
