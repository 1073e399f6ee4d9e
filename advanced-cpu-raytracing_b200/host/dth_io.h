// File readers/writers used by the host mirror: PLY (the reference uses happly, parser.cpp:1404-1416),
// PNG (stb_image / stb_image_write in the reference, LDRImage.h:40, main.cpp:195) and OpenEXR
// (tinyexr, HDRImage.h:51).  Independent implementations; only the subsets our scenes need.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace dth {

struct PlyMesh {
    std::vector<double> positions;      // xyz per vertex, as happly's getVertexPositions() (double)
    std::vector<int> face_counts;       // indices per face
    std::vector<int> face_indices;      // concatenated
    // Streamed path: binary little-endian files whose faces are all triangles are decoded by parallel workers
    // straight into the two flat arrays the scene keeps -- float xyz (the value the reference ends up with: happly widens a float
    // property to double, Mesh narrows it back, parser.cpp:1404-1416 / mesh.cpp:7-13) and int32 index triples -- without the
    // per-property double / per-face vector detour above.  `streamed` tells the caller which pair of arrays is filled.
    bool streamed = false;
    std::vector<float> positions_f32;   // xyz per vertex
    std::vector<int> triangles;         // 3 zero-based vertex indices per face
};
bool ply_load(const std::string& path, PlyMesh& out, std::string& err);
// run fn(begin, end) over [0, n) on up to 16 threads (ranges of at least `grain` items); used by the streamed loader
void parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t)>& fn);

struct ImageData {
    int width = 0, height = 0, channels = 0;
    bool is_hdr = false;
    std::vector<uint8_t> u8;            // LDR: w*h*channels
    std::vector<float> f32;             // HDR: w*h*3
};
bool png_load(const std::string& path, ImageData& out, std::string& err);
bool exr_load(const std::string& path, ImageData& out, std::string& err);   // scanline, NONE/ZIP/ZIPS, half/float
bool png_write(const std::string& path, int w, int h, const uint8_t* rgb, std::string& err);
// dth_output.cpp: band-parallel PNG encoder (same pixels as png_write) and the flat-RGBE .hdr writer
bool png_write_parallel(const std::string& path, int w, int h, const uint8_t* rgb, int n_threads, std::string& err);
bool hdr_write(const std::string& path, int w, int h, const float* rgb, std::string& err);

}  // namespace dth

extern "C" void dth_internal_set_error(const char* msg);   // sets dth_last_error() of the calling thread
