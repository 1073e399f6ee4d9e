#include "dth_io.h"
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <atomic>
#include <thread>

namespace dth {

static bool read_file(const std::string& path, std::vector<uint8_t>& buf) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    size_t r = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return r == (size_t)n;
}

void parallel_for(size_t n, size_t grain, const std::function<void(size_t, size_t)>& fn) {
    if (n == 0) return;
    const size_t hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t nt = std::max<size_t>(1, std::min(hw, n / std::max<size_t>(1, grain)));
    if (nt == 1) { fn(0, n); return; }
    std::vector<std::thread> pool;
    const size_t chunk = (n + nt - 1) / nt;
    for (size_t t = 1; t < nt; t++) pool.emplace_back([&, t] { const size_t b = t * chunk, e = std::min(n, b + chunk); if (b < e) fn(b, e); });
    fn(0, std::min(n, chunk));
    for (auto& th : pool) th.join();
}

// ------------------------------------------------------------------ PLY
namespace {
// read-only view of a whole file: mmap when possible, a heap copy otherwise
struct FileView {
    const uint8_t* p = nullptr; size_t n = 0; bool mapped = false; std::vector<uint8_t> copy;
    bool open(const std::string& path) {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd >= 0) {
            struct stat st;
            if (fstat(fd, &st) == 0 && st.st_size > 0) {
                void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
                if (m != MAP_FAILED) { p = (const uint8_t*)m; n = (size_t)st.st_size; mapped = true; }
            }
            ::close(fd);
        }
        if (mapped) return true;
        if (!read_file(path, copy)) return false;
        p = copy.data(); n = copy.size();
        return true;
    }
    ~FileView() { if (mapped) munmap((void*)p, n); }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
    const uint8_t& operator[](size_t i) const { return p[i]; }
};
enum PlyType { T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64, T_BAD };
PlyType ply_type(const std::string& s) {
    if (s == "char" || s == "int8") return T_I8;
    if (s == "uchar" || s == "uint8") return T_U8;
    if (s == "short" || s == "int16") return T_I16;
    if (s == "ushort" || s == "uint16") return T_U16;
    if (s == "int" || s == "int32") return T_I32;
    if (s == "uint" || s == "uint32") return T_U32;
    if (s == "float" || s == "float32") return T_F32;
    if (s == "double" || s == "float64") return T_F64;
    return T_BAD;
}
int ply_size(PlyType t) {
    switch (t) { case T_I8: case T_U8: return 1; case T_I16: case T_U16: return 2;
                 case T_I32: case T_U32: case T_F32: return 4; case T_F64: return 8; default: return 0; }
}
struct PlyProp { std::string name; bool is_list = false; PlyType count_type = T_BAD, type = T_BAD; };
struct PlyElem { std::string name; size_t count = 0; std::vector<PlyProp> props; };

inline double rd_bin(const uint8_t* p, PlyType t, bool swap) {
    uint8_t b[8];
    int n = ply_size(t);
    if (swap) for (int i = 0; i < n; i++) b[i] = p[n - 1 - i]; else memcpy(b, p, n);
    switch (t) {
        case T_I8: return (double)*(int8_t*)b;
        case T_U8: return (double)*(uint8_t*)b;
        case T_I16: { int16_t v; memcpy(&v, b, 2); return v; }
        case T_U16: { uint16_t v; memcpy(&v, b, 2); return v; }
        case T_I32: { int32_t v; memcpy(&v, b, 4); return v; }
        case T_U32: { uint32_t v; memcpy(&v, b, 4); return v; }
        case T_F32: { float v; memcpy(&v, b, 4); return v; }
        case T_F64: { double v; memcpy(&v, b, 8); return v; }
        default: return 0;
    }
}
}  // namespace

bool ply_load(const std::string& path, PlyMesh& out, std::string& err) {
    FileView buf;                                   // mapped, not copied: a 10 M-triangle PLY is ~190 MB
    if (!buf.open(path)) { err = "cannot read PLY file " + path; return false; }
    // header
    size_t pos = 0;
    auto getline_ = [&](std::string& line) -> bool {
        if (pos >= buf.size()) return false;
        size_t e = pos;
        while (e < buf.size() && buf[e] != '\n') e++;
        line.assign((const char*)&buf[pos], e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e + 1;
        return true;
    };
    std::string line;
    if (!getline_(line) || line.compare(0, 3, "ply") != 0) { err = "not a PLY file: " + path; return false; }
    int fmt = -1;  // 0 ascii, 1 little, 2 big
    std::vector<PlyElem> elems;
    while (getline_(line)) {
        std::istringstream ls(line);
        std::string w;
        ls >> w;
        if (w == "format") {
            std::string f; ls >> f;
            fmt = f == "ascii" ? 0 : f == "binary_little_endian" ? 1 : f == "binary_big_endian" ? 2 : -1;
        } else if (w == "element") {
            PlyElem e; ls >> e.name >> e.count; elems.push_back(e);
        } else if (w == "property") {
            if (elems.empty()) { err = "PLY property before element"; return false; }
            PlyProp p; std::string t; ls >> t;
            if (t == "list") { std::string ct, it; ls >> ct >> it >> p.name; p.is_list = true; p.count_type = ply_type(ct); p.type = ply_type(it); }
            else { p.type = ply_type(t); ls >> p.name; }
            if (p.type == T_BAD) { err = "PLY: unknown property type in '" + line + "'"; return false; }
            elems.back().props.push_back(p);
        } else if (w == "end_header") break;
    }
    if (fmt < 0) { err = "PLY: unknown format"; return false; }
    out.positions.clear(); out.face_counts.clear(); out.face_indices.clear();
    out.streamed = false; out.positions_f32.clear(); out.triangles.clear();

    // ---- streamed path: binary little-endian, elements {vertex (scalars only), face (ONE list of 32-bit indices)}, all triangles ----
    if (fmt == 1 && elems.size() == 2 && elems[0].name == "vertex" && elems[1].name == "face" && elems[1].props.size() == 1 && elems[1].props[0].is_list &&
        ply_size(elems[1].props[0].count_type) == 1 && (elems[1].props[0].type == T_I32 || elems[1].props[0].type == T_U32) &&
        (elems[1].props[0].name == "vertex_indices" || elems[1].props[0].name == "vertex_index")) {
        size_t stride = 0, off[3] = {0, 0, 0}; PlyType ty[3] = {T_BAD, T_BAD, T_BAD};
        bool scalars = true;
        for (auto& p : elems[0].props) {
            if (p.is_list) { scalars = false; break; }
            const int a = p.name == "x" ? 0 : p.name == "y" ? 1 : p.name == "z" ? 2 : -1;
            if (a >= 0) { off[a] = stride; ty[a] = p.type; }
            stride += (size_t)ply_size(p.type);
        }
        const size_t nv = elems[0].count, nf = elems[1].count;
        const bool xyz_ok = (ty[0] == T_F32 || ty[0] == T_F64) && ty[1] == ty[0] && ty[2] == ty[0];
        if (scalars && xyz_ok && buf.size() - pos == nv * stride + nf * 13) {            // 13 = count byte + three 32-bit indices
            const uint8_t* vp = buf.data() + pos;
            const uint8_t* fp = vp + nv * stride;
            std::atomic<bool> all_tris(true), in_range(true);
            out.triangles.resize(nf * 3);
            parallel_for(nf, 1 << 16, [&](size_t b, size_t e) {
                bool tris = true, ok = true;
                for (size_t i = b; i < e; i++) {
                    const uint8_t* r = fp + i * 13;
                    tris &= r[0] == 3;
                    int32_t ix[3]; memcpy(ix, r + 1, 12);
                    ok &= (uint32_t)ix[0] < nv && (uint32_t)ix[1] < nv && (uint32_t)ix[2] < nv;
                    out.triangles[i * 3] = ix[0]; out.triangles[i * 3 + 1] = ix[1]; out.triangles[i * 3 + 2] = ix[2];
                }
                if (!tris) all_tris = false;
                if (!ok) in_range = false;
            });
            if (all_tris && !in_range) { err = "PLY: face index out of range in " + path; return false; }
            if (all_tris) {
                out.positions_f32.resize(nv * 3);
                const bool f32 = ty[0] == T_F32;
                parallel_for(nv, 1 << 16, [&](size_t b, size_t e) {
                    for (size_t i = b; i < e; i++)
                        for (int a = 0; a < 3; a++) {
                            const uint8_t* s = vp + i * stride + off[a];
                            if (f32) memcpy(&out.positions_f32[i * 3 + a], s, 4);
                            else { double d; memcpy(&d, s, 8); out.positions_f32[i * 3 + a] = (float)d; }
                        }
                });
                out.streamed = true;
                return true;
            }
            out.triangles.clear();                     // quads or polygons (sizes happened to add up): the generic reader below
        }
    }

    const bool swap = (fmt == 2);
    // ascii tokenizer state
    const char* ap = (const char*)buf.data() + pos;
    const char* aend = (const char*)buf.data() + buf.size();
    auto next_ascii = [&]() -> double {
        while (ap < aend && isspace((unsigned char)*ap)) ap++;
        char* e = nullptr;
        double v = strtod(ap, &e);
        ap = e ? e : aend;
        return v;
    };
    const uint8_t* bp = buf.data() + pos;
    const uint8_t* bend = buf.data() + buf.size();

    for (auto& el : elems) {
        const bool is_vertex = el.name == "vertex";
        const bool is_face = el.name == "face";
        int ix = -1, iy = -1, iz = -1, ilist = -1;
        for (size_t k = 0; k < el.props.size(); k++) {
            auto& p = el.props[k];
            if (is_vertex && !p.is_list) { if (p.name == "x") ix = (int)k; if (p.name == "y") iy = (int)k; if (p.name == "z") iz = (int)k; }
            if (is_face && p.is_list && (p.name == "vertex_indices" || p.name == "vertex_index")) ilist = (int)k;
        }
        if (is_vertex) {
            if (ix < 0 || iy < 0 || iz < 0) { err = "PLY: vertex element lacks x/y/z"; return false; }
            out.positions.resize(el.count * 3);
        }
        if (is_face) { out.face_counts.reserve(el.count); out.face_indices.reserve(el.count * 3); }
        for (size_t r = 0; r < el.count; r++) {
            for (size_t k = 0; k < el.props.size(); k++) {
                auto& p = el.props[k];
                if (!p.is_list) {
                    double v;
                    if (fmt == 0) v = next_ascii();
                    else { int n = ply_size(p.type); if (bp + n > bend) { err = "PLY: truncated"; return false; } v = rd_bin(bp, p.type, swap); bp += n; }
                    if (is_vertex) { if ((int)k == ix) out.positions[r * 3 + 0] = v; else if ((int)k == iy) out.positions[r * 3 + 1] = v; else if ((int)k == iz) out.positions[r * 3 + 2] = v; }
                } else {
                    int cnt;
                    if (fmt == 0) cnt = (int)next_ascii();
                    else { int n = ply_size(p.count_type); if (bp + n > bend) { err = "PLY: truncated"; return false; } cnt = (int)rd_bin(bp, p.count_type, swap); bp += n; }
                    const bool keep = is_face && (int)k == ilist;
                    if (keep) out.face_counts.push_back(cnt);
                    for (int c = 0; c < cnt; c++) {
                        double v;
                        if (fmt == 0) v = next_ascii();
                        else { int n = ply_size(p.type); if (bp + n > bend) { err = "PLY: truncated"; return false; } v = rd_bin(bp, p.type, swap); bp += n; }
                        if (keep) out.face_indices.push_back((int)v);
                    }
                }
            }
        }
    }
    return true;
}

// ------------------------------------------------------------------ zlib helpers
static bool zinflate(const uint8_t* src, size_t n, std::vector<uint8_t>& dst, size_t expected) {
    dst.resize(expected);
    uLongf dl = (uLongf)expected;
    int r = uncompress(dst.data(), &dl, src, (uLong)n);
    if (r != Z_OK) return false;
    dst.resize(dl);
    return true;
}

// ------------------------------------------------------------------ PNG
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

bool png_load(const std::string& path, ImageData& out, std::string& err) {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf)) { err = "cannot read image " + path; return false; }
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (buf.size() < 8 || memcmp(buf.data(), sig, 8) != 0) { err = "not a PNG: " + path; return false; }
    size_t p = 8;
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    while (p + 8 <= buf.size()) {
        uint32_t len = be32(&buf[p]);
        std::string type((const char*)&buf[p + 4], 4);
        const uint8_t* d = &buf[p + 8];
        if (p + 12 + len > buf.size()) { err = "PNG truncated"; return false; }
        if (type == "IHDR") { w = (int)be32(d); h = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; }
        else if (type == "PLTE") plte.assign(d, d + len);
        else if (type == "tRNS") trns.assign(d, d + len);
        else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
        else if (type == "IEND") break;
        p += 12 + len;
    }
    if (interlace) { err = "interlaced PNG unsupported"; return false; }
    if (depth != 8 && depth != 16 && !(ctype == 3 || ctype == 0)) { err = "PNG bit depth unsupported"; return false; }
    int nch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!nch) { err = "PNG colour type unsupported"; return false; }
    size_t bpp_bits = (size_t)nch * depth;
    size_t stride = ((size_t)w * bpp_bits + 7) / 8;
    size_t bpp = std::max<size_t>(1, bpp_bits / 8);
    std::vector<uint8_t> raw;
    if (!zinflate(idat.data(), idat.size(), raw, (stride + 1) * (size_t)h)) { err = "PNG inflate failed"; return false; }
    if (raw.size() < (stride + 1) * (size_t)h) { err = "PNG data short"; return false; }
    std::vector<uint8_t> img(stride * (size_t)h);
    for (int y = 0; y < h; y++) {
        const uint8_t* in = &raw[(stride + 1) * (size_t)y];
        uint8_t ft = in[0];
        in++;
        uint8_t* cur = &img[stride * (size_t)y];
        const uint8_t* prev = y ? &img[stride * (size_t)(y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            int a = x >= bpp ? cur[x - bpp] : 0;
            int b = prev ? prev[x] : 0;
            int c = (prev && x >= bpp) ? prev[x - bpp] : 0;
            int v = in[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: { int pp = a + b - c; int pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
                          v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: err = "PNG bad filter"; return false;
            }
            cur[x] = (uint8_t)v;
        }
    }
    // expand to 8-bit interleaved with stb_image's channel conventions
    int och = ctype == 3 ? (trns.empty() ? 3 : 4) : nch;
    out.width = w; out.height = h; out.channels = och; out.is_hdr = false;
    out.u8.assign((size_t)w * h * och, 0);
    for (int y = 0; y < h; y++) {
        const uint8_t* row = &img[stride * (size_t)y];
        for (int x = 0; x < w; x++) {
            uint8_t* o = &out.u8[((size_t)y * w + x) * och];
            if (ctype == 3) {
                int idx;
                if (depth == 8) idx = row[x];
                else { int per = 8 / depth; int sh = (per - 1 - (x % per)) * depth; idx = (row[x / per] >> sh) & ((1 << depth) - 1); }
                for (int k = 0; k < 3; k++) o[k] = (size_t)(idx * 3 + k) < plte.size() ? plte[idx * 3 + k] : 0;
                if (och == 4) o[3] = (size_t)idx < trns.size() ? trns[idx] : 255;
            } else if (depth == 8) {
                for (int k = 0; k < nch; k++) o[k] = row[(size_t)x * nch + k];
            } else if (depth == 16) {
                for (int k = 0; k < nch; k++) o[k] = row[((size_t)x * nch + k) * 2];
            } else {  // grey < 8 bit
                int per = 8 / depth; int sh = (per - 1 - (x % per)) * depth;
                int v = (row[x / per] >> sh) & ((1 << depth) - 1);
                o[0] = (uint8_t)(v * 255 / ((1 << depth) - 1));
            }
        }
    }
    return true;
}

bool png_write(const std::string& path, int w, int h, const uint8_t* rgb, std::string& err) {
    std::vector<uint8_t> raw((size_t)(w * 3 + 1) * h);
    for (int y = 0; y < h; y++) {
        raw[(size_t)(w * 3 + 1) * y] = 0;
        memcpy(&raw[(size_t)(w * 3 + 1) * y + 1], rgb + (size_t)w * 3 * y, (size_t)w * 3);
    }
    uLongf cl = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(cl);
    if (compress2(comp.data(), &cl, raw.data(), (uLong)raw.size(), 6) != Z_OK) { err = "PNG deflate failed"; return false; }
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    fwrite(sig, 1, 8, f);
    auto chunk = [&](const char* type, const uint8_t* d, uint32_t n) {
        uint8_t l[4] = {(uint8_t)(n >> 24), (uint8_t)(n >> 16), (uint8_t)(n >> 8), (uint8_t)n};
        fwrite(l, 1, 4, f);
        fwrite(type, 1, 4, f);
        if (n) fwrite(d, 1, n, f);
        uLong c = crc32(0L, (const Bytef*)type, 4);
        if (n) c = crc32(c, d, n);
        uint8_t cb[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
        fwrite(cb, 1, 4, f);
    };
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                        (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)cl);
    chunk("IEND", nullptr, 0);
    fclose(f);
    return true;
}

// ------------------------------------------------------------------ EXR (scanline; NONE / ZIPS / ZIP)
static float half_to_float(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023, u;
    if (e == 0) {
        if (m == 0) u = s << 31;
        else { int k = 0; while (!(m & 1024)) { m <<= 1; k++; } m &= 1023; u = (s << 31) | ((uint32_t)(127 - 15 - k + 1) << 23) | (m << 13); }
    } else if (e == 31) u = (s << 31) | 0x7f800000u | (m << 13);
    else u = (s << 31) | ((e - 15 + 127) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}

bool exr_load(const std::string& path, ImageData& out, std::string& err) {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf)) { err = "cannot read EXR " + path; return false; }
    if (buf.size() < 8 || buf[0] != 0x76 || buf[1] != 0x2f || buf[2] != 0x31 || buf[3] != 0x01) { err = "not an EXR: " + path; return false; }
    uint32_t ver; memcpy(&ver, &buf[4], 4);
    if (ver & 0x200) { err = "tiled EXR unsupported"; return false; }
    if (ver & 0x1800) { err = "multipart/deep EXR unsupported"; return false; }
    size_t p = 8;
    struct Ch { std::string name; int type; };
    std::vector<Ch> chans;
    int comp = 0, xmin = 0, ymin = 0, xmax = -1, ymax = -1;
    auto cstr = [&](std::string& s) { size_t e = p; while (e < buf.size() && buf[e]) e++; s.assign((const char*)&buf[p], e - p); p = e + 1; };
    for (;;) {
        if (p >= buf.size()) { err = "EXR header truncated"; return false; }
        if (buf[p] == 0) { p++; break; }
        std::string an, at; cstr(an); cstr(at);
        int32_t sz; memcpy(&sz, &buf[p], 4); p += 4;
        const uint8_t* v = &buf[p];
        if (an == "channels") {
            size_t q = 0;
            while (q < (size_t)sz && v[q]) {
                Ch c; size_t e = q; while (v[e]) e++; c.name.assign((const char*)&v[q], e - q); q = e + 1;
                int32_t t; memcpy(&t, &v[q], 4); c.type = t; q += 16;
                chans.push_back(c);
            }
        } else if (an == "compression") comp = v[0];
        else if (an == "dataWindow") { int32_t b[4]; memcpy(b, v, 16); xmin = b[0]; ymin = b[1]; xmax = b[2]; ymax = b[3]; }
        p += (size_t)sz;
    }
    if (comp != 0 && comp != 2 && comp != 3) { err = "EXR compression unsupported (only NONE/ZIPS/ZIP)"; return false; }
    int w = xmax - xmin + 1, h = ymax - ymin + 1;
    if (w <= 0 || h <= 0) { err = "EXR bad dataWindow"; return false; }
    int lines_per_block = comp == 3 ? 16 : 1;
    int nblocks = (h + lines_per_block - 1) / lines_per_block;
    size_t line_bytes = 0;
    for (auto& c : chans) line_bytes += (size_t)w * (c.type == 1 ? 2 : 4);
    int ir = -1, ig = -1, ib = -1, iy = -1;
    for (size_t k = 0; k < chans.size(); k++) { if (chans[k].name == "R") ir = (int)k; if (chans[k].name == "G") ig = (int)k; if (chans[k].name == "B") ib = (int)k; if (chans[k].name == "Y") iy = (int)k; }
    if (ir < 0 && iy >= 0) ir = ig = ib = iy;
    if (ir < 0 || ig < 0 || ib < 0) { err = "EXR lacks R/G/B channels"; return false; }
    out.width = w; out.height = h; out.channels = 3; out.is_hdr = true;
    out.f32.assign((size_t)w * h * 3, 0.f);
    if (p + 8 * (size_t)nblocks > buf.size()) { err = "EXR offsets truncated"; return false; }
    std::vector<uint8_t> tmp, un;
    for (int b = 0; b < nblocks; b++) {
        uint64_t off; memcpy(&off, &buf[p + 8 * (size_t)b], 8);
        if (off + 8 > buf.size()) { err = "EXR chunk offset out of range"; return false; }
        int32_t y, sz; memcpy(&y, &buf[off], 4); memcpy(&sz, &buf[off + 4], 4);
        const uint8_t* d = &buf[off + 8];
        int nl = std::min(lines_per_block, ymax - y + 1);
        size_t expect = line_bytes * (size_t)nl;
        const uint8_t* data = d;
        if (comp != 0 && (size_t)sz < expect) {
            if (!zinflate(d, (size_t)sz, tmp, expect) || tmp.size() != expect) { err = "EXR inflate failed"; return false; }
            // predictor
            for (size_t k = 1; k < expect; k++) tmp[k] = (uint8_t)(tmp[k - 1] + tmp[k] - 128);
            // de-interleave
            un.resize(expect);
            size_t half = (expect + 1) / 2;
            for (size_t k = 0; k < expect; k++) un[k] = (k & 1) ? tmp[half + k / 2] : tmp[k / 2];
            data = un.data();
        }
        for (int l = 0; l < nl; l++) {
            const uint8_t* lp = data + line_bytes * (size_t)l;
            int row = y - ymin + l;
            size_t coff = 0;
            for (size_t k = 0; k < chans.size(); k++) {
                size_t bs = chans[k].type == 1 ? 2 : 4;
                for (int c3 = 0; c3 < 3; c3++) {
                    int want = c3 == 0 ? ir : c3 == 1 ? ig : ib;
                    if ((int)k != want) continue;
                    for (int x = 0; x < w; x++) {
                        float f;
                        if (chans[k].type == 1) { uint16_t hv; memcpy(&hv, lp + coff + 2 * (size_t)x, 2); f = half_to_float(hv); }
                        else if (chans[k].type == 2) memcpy(&f, lp + coff + 4 * (size_t)x, 4);
                        else { uint32_t u; memcpy(&u, lp + coff + 4 * (size_t)x, 4); f = (float)u; }
                        out.f32[((size_t)row * w + x) * 3 + c3] = f;
                    }
                }
                coff += bs * (size_t)w;
            }
        }
    }
    return true;
}

}  // namespace dth
