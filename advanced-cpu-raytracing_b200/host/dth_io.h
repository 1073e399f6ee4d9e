// File readers/writers used by the host mirror: PLY (the reference uses happly, parser.cpp:1404-1416),
// PNG (stb_image / stb_image_write in the reference, LDRImage.h:40, main.cpp:195) and OpenEXR
// (tinyexr, HDRImage.h:51).  Independent implementations; only the subsets our scenes need.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dth {

struct PlyMesh {
    std::vector<double> positions;      // xyz per vertex, as happly's getVertexPositions() (double)
    std::vector<int> face_counts;       // indices per face
    std::vector<int> face_indices;      // concatenated
};
bool ply_load(const std::string& path, PlyMesh& out, std::string& err);

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
