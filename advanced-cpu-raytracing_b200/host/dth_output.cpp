// Asynchronous image output.  The reference encodes on the main thread, inside its timed region:
// stbi_write_hdr + stbi_write_png after every camera (main.cpp:186-195; stbi_zlib_compress is ~0.09 s per 1440x720 frame in
// its gprof run).  With frames rendered in milliseconds the encode IS the frame time, so here it leaves the render thread:
//   * dth_writer: a queue + worker threads; submit() copies the pixels (the caller reuses its frame buffers for the next
//     camera at once) and returns, wait() joins the outstanding files and reports the first error;
//   * the PNG encoder itself is parallel: the rows are cut into bands, every band is deflated on its own (raw deflate, ended
//     with a sync flush so that it stops on a byte boundary) and the bands are concatenated into ONE zlib stream whose Adler-32
//     is combined from the bands' checksums -- the image a decoder sees is byte-identical to the serial encoder's.
// Independent implementation on zlib's public API; no stb / tinyexr code.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

#include "../../include/dorktracer_host.h"
#include "dth_io.h"

namespace dth {

namespace {

struct Band { std::vector<uint8_t> z; uLong adler = 1; size_t raw_len = 0; bool ok = true; };

// rows [y0, y1) of an RGB image as PNG scanlines (filter byte 0 + 3w bytes), raw-deflated; `last` ends the deflate stream
void deflate_band(const uint8_t* rgb, int w, int y0, int y1, bool last, int level, Band& out) {
    const size_t row = (size_t)w * 3, n = (row + 1) * (size_t)(y1 - y0);
    std::vector<uint8_t> raw(n);
    for (int y = y0; y < y1; y++) {
        uint8_t* d = &raw[(row + 1) * (size_t)(y - y0)];
        d[0] = 0;
        memcpy(d + 1, rgb + row * (size_t)y, row);
    }
    out.raw_len = n;
    out.adler = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)n);
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { out.ok = false; return; }
    out.z.resize(deflateBound(&zs, (uLong)n) + 16);
    zs.next_in = raw.data(); zs.avail_in = (uInt)n;
    zs.next_out = out.z.data(); zs.avail_out = (uInt)out.z.size();
    const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
    out.ok = last ? rc == Z_STREAM_END : (rc == Z_OK && zs.avail_in == 0);
    out.z.resize(out.z.size() - zs.avail_out);
    deflateEnd(&zs);
}

void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

bool write_chunk(FILE* f, const char* type, const uint8_t* d, size_t n) {
    uint8_t l[4]; put_be32(l, (uint32_t)n);
    uLong c = crc32(0L, (const Bytef*)type, 4);
    if (n) c = crc32(c, d, (uInt)n);
    uint8_t cb[4]; put_be32(cb, (uint32_t)c);
    return fwrite(l, 1, 4, f) == 4 && fwrite(type, 1, 4, f) == 4 && (n == 0 || fwrite(d, 1, n, f) == n) && fwrite(cb, 1, 4, f) == 4;
}

}  // namespace

// PNG (8-bit RGB, filter 0) with the rows deflated in `n_threads` parallel bands (0 = hardware concurrency, at most 16).
bool png_write_parallel(const std::string& path, int w, int h, const uint8_t* rgb, int n_threads, std::string& err) {
    if (w <= 0 || h <= 0 || !rgb) { err = "png_write_parallel: empty image"; return false; }
    if (n_threads <= 0) n_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    const int rows_per_band = std::max(16, (h + n_threads * 2 - 1) / (n_threads * 2));       // two bands per thread: the bands' costs differ with content
    const int n_bands = (h + rows_per_band - 1) / rows_per_band;
    std::vector<Band> bands((size_t)n_bands);
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= n_bands) break;
            deflate_band(rgb, w, b * rows_per_band, std::min(h, (b + 1) * rows_per_band), b == n_bands - 1, 6, bands[(size_t)b]);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::min(n_threads, n_bands); t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    size_t total = 2 + 4;
    uLong adler = adler32(0L, Z_NULL, 0);
    for (auto& b : bands) {
        if (!b.ok) { err = "PNG deflate failed"; return false; }
        total += b.z.size();
        adler = adler32_combine(adler, b.adler, (z_off_t)b.raw_len);
    }
    std::vector<uint8_t> idat; idat.reserve(total);
    idat.push_back(0x78); idat.push_back(0x9C);                               // zlib header: deflate, 32 K window, default level
    for (auto& b : bands) idat.insert(idat.end(), b.z.begin(), b.z.end());
    uint8_t a4[4]; put_be32(a4, (uint32_t)adler);
    idat.insert(idat.end(), a4, a4 + 4);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    uint8_t ihdr[13]; put_be32(ihdr, (uint32_t)w); put_be32(ihdr + 4, (uint32_t)h); ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = ihdr[11] = ihdr[12] = 0;
    bool ok = fwrite(sig, 1, 8, f) == 8 && write_chunk(f, "IHDR", ihdr, 13) && write_chunk(f, "IDAT", idat.data(), idat.size()) && write_chunk(f, "IEND", nullptr, 0);
    ok = (fclose(f) == 0) && ok;
    if (!ok) err = "short write to " + path;
    return ok;
}

// Radiance .hdr, flat (non-RLE) RGBE scanlines: what stbi_write_hdr produces for the reference's HDR output (main.cpp:191).
bool hdr_write(const std::string& path, int w, int h, const float* rgb, std::string& err) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", h, w);
    std::vector<unsigned char> row((size_t)w * 4);
    bool ok = true;
    for (int y = 0; y < h && ok; y++) {
        for (int x = 0; x < w; x++) {
            const float* p = rgb + ((size_t)y * w + x) * 3;
            float m = p[0] > p[1] ? p[0] : p[1]; if (p[2] > m) m = p[2];
            unsigned char* o = &row[(size_t)x * 4];
            if (!(m > 1e-32f)) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
            int e; const float n = frexpf(m, &e) * 256.0f / m;
            o[0] = (unsigned char)(p[0] * n); o[1] = (unsigned char)(p[1] * n); o[2] = (unsigned char)(p[2] * n); o[3] = (unsigned char)(e + 128);
        }
        ok = fwrite(row.data(), 1, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) err = "short write to " + path;
    return ok;
}

}  // namespace dth

// ------------------------------------------------------------------ C ABI: the writer queue
struct dth_writer {
    struct Job { std::string path; int w = 0, h = 0; std::vector<uint8_t> ldr; std::vector<float> hdr; };
    std::mutex mu;
    std::condition_variable cv_work, cv_idle;
    std::deque<Job> jobs;
    int in_flight = 0;
    bool stop = false;
    int encode_threads = 0;
    std::string first_error;
    double busy_seconds = 0.0;
    std::vector<std::thread> workers;

    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !jobs.empty(); });
                if (jobs.empty()) return;
                j = std::move(jobs.front()); jobs.pop_front();
            }
            const auto t0 = std::chrono::steady_clock::now();
            std::string err;
            const bool ok = j.hdr.empty() ? dth::png_write_parallel(j.path, j.w, j.h, j.ldr.data(), encode_threads, err)
                                          : dth::hdr_write(j.path, j.w, j.h, j.hdr.data(), err);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (!ok && first_error.empty()) first_error = err;
                busy_seconds += dt;
                in_flight--;
            }
            cv_idle.notify_all();
        }
    }
};

extern "C" {

dth_writer* dth_writer_create(int n_files_in_parallel, int encode_threads) {
    dth_writer* w = new dth_writer();
    w->encode_threads = encode_threads;
    const int n = std::max(1, std::min(8, n_files_in_parallel));
    for (int i = 0; i < n; i++) w->workers.emplace_back([w] { w->run(); });
    return w;
}

static int submit(dth_writer* w, dth_writer::Job&& j) {
    if (!w) return DT_ERR_INVALID;
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->jobs.push_back(std::move(j));
        w->in_flight++;
    }
    w->cv_work.notify_one();
    return DT_OK;
}

int dth_writer_submit_png(dth_writer* w, const char* path, int width, int height, const uint8_t* rgb) {
    if (!path || !rgb || width <= 0 || height <= 0) return DT_ERR_INVALID;
    dth_writer::Job j; j.path = path; j.w = width; j.h = height;
    j.ldr.assign(rgb, rgb + (size_t)width * height * 3);
    return submit(w, std::move(j));
}

int dth_writer_submit_hdr(dth_writer* w, const char* path, int width, int height, const float* rgb) {
    if (!path || !rgb || width <= 0 || height <= 0) return DT_ERR_INVALID;
    dth_writer::Job j; j.path = path; j.w = width; j.h = height;
    j.hdr.assign(rgb, rgb + (size_t)width * height * 3);
    return submit(w, std::move(j));
}

int dth_writer_wait(dth_writer* w, double* busy_seconds) {
    if (!w) return DT_ERR_INVALID;
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv_idle.wait(lk, [&] { return w->in_flight == 0; });
    if (busy_seconds) *busy_seconds = w->busy_seconds;
    if (!w->first_error.empty()) { dth_internal_set_error(w->first_error.c_str()); w->first_error.clear(); return DT_ERR_INVALID; }
    return DT_OK;
}

void dth_writer_destroy(dth_writer* w) {
    if (!w) return;
    {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv_idle.wait(lk, [&] { return w->in_flight == 0; });
        w->stop = true;
    }
    w->cv_work.notify_all();
    for (auto& t : w->workers) t.join();
    delete w;
}

int dth_write_png_parallel(const char* path, int width, int height, const uint8_t* rgb, int n_threads) {
    std::string err;
    if (!path || !dth::png_write_parallel(path, width, height, rgb, n_threads, err)) { dth_internal_set_error(err.c_str()); return DT_ERR_INVALID; }
    return DT_OK;
}

int dth_write_hdr(const char* path, int width, int height, const float* rgb) {
    std::string err;
    if (!path || !rgb || !dth::hdr_write(path, width, height, rgb, err)) { dth_internal_set_error(err.c_str()); return DT_ERR_INVALID; }
    return DT_OK;
}

}  // extern "C"
