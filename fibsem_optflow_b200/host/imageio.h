// imageio.h -- the image I/O the job driver needs, without OpenCV / libpng / libtiff.
//
// Stands in for cv::imread(..., IMREAD_GRAYSCALE) (reference src/optflow.cpp:106,119) and the float
// cv::imwrite of the flow planes (:478-484); the prescale cv::resize of :111,124 runs on the device
// (tvl1_prescale_u8).  Readers: PNG (8/16-bit grey, grey+alpha, RGB(A), palette; plain or Adam7-interlaced)
// through zlib, binary PGM (P5), and 8/16-bit grey TIFF in strips -- raw, PackBits, LZW or Deflate,
// with or without the horizontal predictor (cv::imwrite's own TIFFs are LZW + predictor).  Colour is
// reduced to grey and 16 bits to 8 the way cv::imread does it; every path is pinned against cv2
// (tests/test_host_tools.py).  Writer: uncompressed 32-bit float single-strip TIFF, which is what
// support_scripts/upload_matches.py opens with PIL (upload_matches.py:34-37).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace imio {

struct Gray8 {
    int w = 0, h = 0;
    std::vector<uint8_t> px;   // tight rows
};

inline bool read_file(const std::string& path, std::vector<uint8_t>& out)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == out.size();
}

// ---- PNG -----------------------------------------------------------------------------------

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline uint8_t rgb_to_gray(int r, int g, int b)
{
    // libpng png_set_rgb_to_gray(0.299, 0.587), the conversion OpenCV's PNG reader asks for when
    // cv::imread is given IMREAD_GRAYSCALE: 15-bit fixed point with the coefficients libpng derives
    // (29900 * 32768 / 100000 = 9797, 58700 * 32768 / 100000 = 19234, blue = the rest) and -- the
    // historical libpng 1.6 path without a gamma table -- truncation.  Pinned against cv2.imread
    // (tests/test_host_tools.py).
    return (uint8_t)((9797 * r + 19234 * g + 3737 * b) >> 15);
}

inline bool decode_png(const std::vector<uint8_t>& f, Gray8& img, std::string& err)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (f.size() < 33 || std::memcmp(f.data(), sig, 8) != 0) { err = "not a PNG"; return false; }
    size_t p = 8;
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    while (p + 12 <= f.size()) {
        const uint32_t len = be32(&f[p]);
        const char* tag = (const char*)&f[p + 4];
        if (p + 12 + len > f.size()) { err = "truncated PNG chunk"; return false; }
        const uint8_t* d = &f[p + 8];
        if (!std::memcmp(tag, "IHDR", 4)) {
            w = (int)be32(d); h = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
        } else if (!std::memcmp(tag, "PLTE", 4)) {
            plte.assign(d, d + len);
        } else if (!std::memcmp(tag, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!std::memcmp(tag, "IEND", 4)) {
            break;
        }
        p += 12 + len;
    }
    if (w <= 0 || h <= 0) { err = "PNG without IHDR"; return false; }
    if (interlace > 1) { err = "bad PNG interlace method"; return false; }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "bad PNG colour type"; return false;
    }
    if (!(depth == 8 || depth == 16 || (depth < 8 && (ctype == 0 || ctype == 3)))) { err = "unsupported PNG bit depth"; return false; }
    const size_t bpp_bits = (size_t)channels * depth;
    const size_t bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;
    // the sub-images of the stream: the whole image, or the seven Adam7 passes
    struct Pass { int x0, y0, dx, dy, pw, ph; size_t stride; };
    std::vector<Pass> passes;
    if (!interlace) {
        passes.push_back({0, 0, 1, 1, w, h, ((size_t)w * bpp_bits + 7) / 8});
    } else {
        static const int X0[7] = {0, 4, 0, 2, 0, 1, 0}, Y0[7] = {0, 0, 4, 0, 2, 0, 1};
        static const int DX[7] = {8, 8, 4, 4, 2, 2, 1}, DY[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int k = 0; k < 7; k++) {
            const int pw = (w - X0[k] + DX[k] - 1) / DX[k], ph = (h - Y0[k] + DY[k] - 1) / DY[k];
            if (pw > 0 && ph > 0) passes.push_back({X0[k], Y0[k], DX[k], DY[k], pw, ph, ((size_t)pw * bpp_bits + 7) / 8});
        }
    }
    size_t total = 0;
    for (const Pass& ps : passes) total += (ps.stride + 1) * (size_t)ps.ph;
    std::vector<uint8_t> raw(total);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) {
        err = "PNG inflate failed";
        return false;
    }
    img.w = w; img.h = h;
    img.px.resize((size_t)w * h);
    // one pixel of an unfiltered scanline -> 8-bit grey, the way cv::imread(IMREAD_GRAYSCALE) reduces it
    auto to_gray = [&](const uint8_t* c, int x, uint8_t& out) -> bool {
        if (ctype == 0 || ctype == 4) {
            if (depth == 8) out = c[(size_t)x * channels];
            else if (depth == 16) out = c[(size_t)x * channels * 2];   // high byte, like libpng's strip_16
            else {
                const int per = 8 / depth, sh = (per - 1 - x % per) * depth;
                const int v = (c[x / per] >> sh) & ((1 << depth) - 1);
                out = (uint8_t)(v * 255 / ((1 << depth) - 1));
            }
        } else if (ctype == 3) {
            int idx;
            if (depth == 8) idx = c[x];
            else { const int per = 8 / depth, sh = (per - 1 - x % per) * depth; idx = (c[x / per] >> sh) & ((1 << depth) - 1); }
            if ((size_t)idx * 3 + 2 >= plte.size()) return false;
            out = rgb_to_gray(plte[idx * 3], plte[idx * 3 + 1], plte[idx * 3 + 2]);
        } else {
            const size_t step = depth == 16 ? 2 : 1;
            const uint8_t* q = c + (size_t)x * channels * step;
            out = rgb_to_gray(q[0], q[step], q[2 * step]);
        }
        return true;
    };
    size_t base = 0;
    for (const Pass& ps : passes) {
        const size_t stride = ps.stride;
        const std::vector<uint8_t> zeros(stride, 0);
        const uint8_t* prev = zeros.data();
        for (int y = 0; y < ps.ph; y++) {
            uint8_t* row = &raw[base + (stride + 1) * (size_t)y];
            const int ft = row[0];
            uint8_t* c = row + 1;
            // unfilter in place: one tight loop per filter type (the decode of a FIB-SEM slice, not its solve, sets
            // the pace of a job: an 8-bit grey 4096^2 PNG is 17 MB of mostly Paeth-filtered bytes)
            const uint8_t* pv = prev;   // the previous scanline of this pass, unfiltered (zeros for the first)
            switch (ft) {
                case 0: break;
                case 1:
                    for (size_t i = bpp; i < stride; i++) c[i] = (uint8_t)(c[i] + c[i - bpp]);
                    break;
                case 2:
                    for (size_t i = 0; i < stride; i++) c[i] = (uint8_t)(c[i] + pv[i]);
                    break;
                case 3:
                    for (size_t i = 0; i < bpp && i < stride; i++) c[i] = (uint8_t)(c[i] + (pv[i] >> 1));
                    for (size_t i = bpp; i < stride; i++) c[i] = (uint8_t)(c[i] + ((c[i - bpp] + pv[i]) >> 1));
                    break;
                case 4: {
                    for (size_t i = 0; i < bpp && i < stride; i++) c[i] = (uint8_t)(c[i] + pv[i]);   // a = c = 0: predictor is b
                    if (bpp == 1) {
                        int a = stride > 0 ? c[0] : 0, cc = stride > 0 ? pv[0] : 0;
                        for (size_t i = 1; i < stride; i++) {
                            const int b = pv[i];
                            const int pa = b - cc, pb = a - cc;                  // p - a, p - b with p = a + b - c
                            const int apa = pa < 0 ? -pa : pa, apb = pb < 0 ? -pb : pb, apc = (pa + pb) < 0 ? -(pa + pb) : (pa + pb);
                            const int pred = (apa <= apb && apa <= apc) ? a : (apb <= apc ? b : cc);
                            a = (uint8_t)(c[i] + pred);
                            c[i] = (uint8_t)a;
                            cc = b;
                        }
                    } else {
                        for (size_t i = bpp; i < stride; i++) {
                            const int a = c[i - bpp], b = pv[i], cc = pv[i - bpp];
                            const int pa = std::abs(b - cc), pb = std::abs(a - cc), pc = std::abs(a + b - 2 * cc);
                            c[i] = (uint8_t)(c[i] + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : cc)));
                        }
                    }
                    break;
                }
                default: err = "bad PNG filter"; return false;
            }
            prev = c;   // stays valid: `raw` is not touched again above this row
            uint8_t* o = &img.px[(size_t)(ps.y0 + y * ps.dy) * w + ps.x0];
            if (ctype == 0 && depth == 8 && ps.dx == 1) {
                std::memcpy(o, c, (size_t)ps.pw);   // 8-bit grey, not interlaced: the scanline is the output row
            } else {
                for (int x = 0; x < ps.pw; x++)
                    if (!to_gray(c, x, o[(size_t)x * ps.dx])) { err = "PNG palette index out of range"; return false; }
            }
        }
        base += (stride + 1) * (size_t)ps.ph;
    }
    return true;
}

// ---- PGM -----------------------------------------------------------------------------------

inline bool decode_pgm(const std::vector<uint8_t>& f, Gray8& img, std::string& err)
{
    size_t p = 2;
    auto token = [&]() -> long {
        for (;;) {
            while (p < f.size() && isspace(f[p])) p++;
            if (p < f.size() && f[p] == '#') { while (p < f.size() && f[p] != '\n') p++; continue; }
            break;
        }
        long v = 0;
        bool any = false;
        while (p < f.size() && isdigit(f[p])) { v = v * 10 + (f[p++] - '0'); any = true; }
        return any ? v : -1;
    };
    const long w = token(), h = token(), mx = token();
    if (w <= 0 || h <= 0 || mx <= 0 || mx > 255) { err = "unsupported PGM header"; return false; }
    p++;   // single whitespace after maxval
    if (p + (size_t)w * h > f.size()) { err = "truncated PGM"; return false; }
    img.w = (int)w; img.h = (int)h;
    img.px.assign(f.begin() + p, f.begin() + p + (size_t)w * h);
    return true;
}

// ---- TIFF: 8- or 16-bit grey in strips; uncompressed, PackBits, LZW or Deflate; horizontal predictor.
// (cv::imwrite's own TIFFs are LZW + predictor 2.)  16-bit samples keep their high byte, which is what
// cv::imread(IMREAD_GRAYSCALE) returns for them.

// TIFF-flavoured LZW: MSB-first codes of 9..12 bits, ClearCode 256, EOI 257, "early change".
// A table entry is a position and a length in the OUTPUT: the string of a new code is the previous string plus one
// byte, and those bytes were just written, one behind the other -- so a code is decoded by one forward copy and no
// chain of prefixes is ever walked (a 4096^2 cv::imwrite TIFF of band-limited noise: 0.32 -> 0.19 s, libtiff: 0.20).
inline bool tiff_lzw(const uint8_t* in, size_t n, std::vector<uint8_t>& out, size_t want)
{
    const size_t base = out.size();
    out.resize(base + want + 8);                 // decoded in place; trimmed at the end
    uint8_t* const o = out.data() + base;
    uint32_t off[4096];
    uint16_t len[4096];
    int next = 258, width = 9, prev = -1;
    size_t prev_pos = 0, prev_len = 0, pos = 0, ip = 0;
    uint64_t acc = 0;
    int bits = 0;
    bool ok = true;
    while (pos < want) {
        while (bits <= 56 && ip < n) { acc = (acc << 8) | in[ip++]; bits += 8; }
        if (bits < width) break;
        const int code = (int)((acc >> (bits - width)) & ((1u << width) - 1));
        bits -= width;
        if (code == 257) break;
        if (code == 256) { next = 258; width = 9; prev = -1; continue; }
        const size_t cur_pos = pos;
        size_t cur_len;
        if (code < 256) {
            o[pos] = (uint8_t)code;
            cur_len = 1;
        } else if (code < next && prev >= 0) {
            cur_len = len[code];
            if (pos + cur_len > want + 8) cur_len = want + 8 - pos;
            const uint8_t* src = o + off[code];
            for (size_t k = 0; k < cur_len; k++) o[pos + k] = src[k];   // forward: source and target may overlap
        } else if (code == next && prev >= 0 && next < 4096) {
            cur_len = prev_len + 1;
            if (pos + cur_len > want + 8) cur_len = want + 8 - pos;   // the strip's last string may run past what is wanted
            const uint8_t* src = o + prev_pos;
            for (size_t k = 0; k < cur_len; k++) o[pos + k] = k < prev_len ? src[k] : src[0];
        } else {
            ok = false;
            break;
        }
        if (prev >= 0 && next < 4096) {
            // string(next) = string(prev) + first byte of string(code): exactly the bytes at prev_pos .. prev_pos + prev_len
            off[next] = (uint32_t)prev_pos;
            len[next] = (uint16_t)(prev_len + 1 > 65535 ? 65535 : prev_len + 1);
            next++;
        }
        pos += cur_len;
        prev = code; prev_pos = cur_pos; prev_len = cur_len;
        if (next + 1 >= (1 << width) && width < 12) width++;   // early change
    }
    out.resize(base + (pos < want ? pos : want));
    return ok && pos >= want;
}

inline bool tiff_packbits(const uint8_t* in, size_t n, std::vector<uint8_t>& out, size_t want)
{
    size_t p = 0;
    while (p < n && out.size() < want) {
        const int c = (int8_t)in[p++];
        if (c >= 0) {
            if (p + (size_t)c + 1 > n) return false;
            out.insert(out.end(), in + p, in + p + c + 1);
            p += (size_t)c + 1;
        } else if (c != -128) {
            if (p >= n) return false;
            out.insert(out.end(), (size_t)(1 - c), in[p++]);
        }
    }
    return out.size() >= want;
}

inline bool decode_tiff(const std::vector<uint8_t>& f, Gray8& img, std::string& err)
{
    if (f.size() < 8) { err = "truncated TIFF"; return false; }
    const bool le = f[0] == 'I';
    auto u16 = [&](size_t o) -> uint32_t { return le ? (f[o] | (f[o + 1] << 8)) : ((f[o] << 8) | f[o + 1]); };
    auto u32 = [&](size_t o) -> uint32_t {
        return le ? (f[o] | (f[o + 1] << 8) | (f[o + 2] << 16) | ((uint32_t)f[o + 3] << 24))
                  : (((uint32_t)f[o] << 24) | (f[o + 1] << 16) | (f[o + 2] << 8) | f[o + 3]);
    };
    if (u16(2) != 42) { err = "not a TIFF"; return false; }
    size_t ifd = u32(4);
    if (ifd + 2 > f.size()) { err = "bad TIFF IFD"; return false; }
    const uint32_t n = u16(ifd);
    uint32_t w = 0, h = 0, bits = 1, spp = 1, comp = 1, photo = 1, rps = 0xffffffffu, pred = 1, tiled = 0;
    std::vector<uint32_t> offs, counts;
    for (uint32_t k = 0; k < n; k++) {
        const size_t e = ifd + 2 + 12 * (size_t)k;
        if (e + 12 > f.size()) { err = "bad TIFF IFD"; return false; }
        const uint32_t tag = u16(e), type = u16(e + 2), cnt = u32(e + 4);
        const size_t esz = type == 3 ? 2 : (type == 4 ? 4 : 1);
        const size_t vo = cnt * esz <= 4 ? e + 8 : u32(e + 8);
        if (vo + cnt * esz > f.size()) { err = "bad TIFF IFD"; return false; }
        auto val = [&](uint32_t i) -> uint32_t { return type == 3 ? u16(vo + 2 * i) : (type == 4 ? u32(vo + 4 * i) : f[vo + i]); };
        switch (tag) {
            case 256: w = val(0); break;
            case 257: h = val(0); break;
            case 258: bits = val(0); break;
            case 259: comp = val(0); break;
            case 262: photo = val(0); break;
            case 277: spp = val(0); break;
            case 278: rps = val(0); break;
            case 317: pred = val(0); break;
            case 322: tiled = 1; break;
            case 273: for (uint32_t i = 0; i < cnt; i++) offs.push_back(val(i)); break;
            case 279: for (uint32_t i = 0; i < cnt; i++) counts.push_back(val(i)); break;
            default: break;
        }
    }
    const bool comp_ok = comp == 1 || comp == 5 || comp == 8 || comp == 32946 || comp == 32773;
    if (!w || !h || (bits != 8 && bits != 16) || spp != 1 || !comp_ok || pred > 2 || tiled || photo > 1 || offs.empty() ||
        counts.size() != offs.size()) {
        err = "only 8/16-bit single-channel TIFF in strips (raw, PackBits, LZW, Deflate) is supported";
        return false;
    }
    if (rps > h) rps = h;
    const size_t bps = bits / 8, rowbytes = (size_t)w * bps;
    img.w = (int)w; img.h = (int)h;
    img.px.resize((size_t)w * h);
    size_t row = 0;
    std::vector<uint8_t> buf;
    for (size_t s = 0; s < offs.size() && row < h; s++) {
        const size_t rows = std::min<size_t>(rps, h - row), bytes = rows * rowbytes;
        if ((size_t)offs[s] + counts[s] > f.size()) { err = "truncated TIFF strip"; return false; }
        const uint8_t* src = &f[offs[s]];
        buf.clear();
        bool ok = true;
        if (comp == 1) {
            if (counts[s] < bytes) ok = false; else buf.assign(src, src + bytes);
        } else if (comp == 5) {
            buf.reserve(bytes);
            ok = tiff_lzw(src, counts[s], buf, bytes);
        } else if (comp == 32773) {
            buf.reserve(bytes);
            ok = tiff_packbits(src, counts[s], buf, bytes);
        } else {
            buf.resize(bytes);
            uLongf got = (uLongf)bytes;
            ok = uncompress(buf.data(), &got, src, (uLong)counts[s]) == Z_OK && got >= bytes;
        }
        if (!ok) { err = "corrupt TIFF strip"; return false; }
        for (size_t r = 0; r < rows; r++) {
            uint8_t* line = &buf[r * rowbytes];
            uint8_t* o = &img.px[(row + r) * w];
            if (bits == 8) {
                if (pred == 2) for (size_t x = 1; x < w; x++) line[x] = (uint8_t)(line[x] + line[x - 1]);
                std::memcpy(o, line, w);
            } else {
                uint16_t acc = 0;
                for (size_t x = 0; x < w; x++) {
                    uint16_t v = le ? (uint16_t)(line[2 * x] | (line[2 * x + 1] << 8)) : (uint16_t)((line[2 * x] << 8) | line[2 * x + 1]);
                    if (pred == 2) { acc = (uint16_t)(acc + v); v = acc; }
                    o[x] = (uint8_t)(v >> 8);
                }
            }
        }
        row += rows;
    }
    if (row < h) { err = "TIFF strips do not cover the image"; return false; }
    if (photo == 0) for (auto& v : img.px) v = (uint8_t)(255 - v);   // WhiteIsZero
    return true;
}

inline bool read_gray8(const std::string& path, Gray8& img, std::string& err)
{
    std::vector<uint8_t> f;
    if (!read_file(path, f) || f.size() < 4) { err = "cannot read " + path; return false; }
    if (f[0] == 0x89 && f[1] == 'P') return decode_png(f, img, err);
    if (f[0] == 'P' && f[1] == '5') return decode_pgm(f, img, err);
    if ((f[0] == 'I' && f[1] == 'I') || (f[0] == 'M' && f[1] == 'M')) return decode_tiff(f, img, err);
    err = "unknown image format: " + path;
    return false;
}


// ---- float TIFF writer -----------------------------------------------------------------------

inline bool write_tiff_f32(const std::string& path, const float* data, int w, int h)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint32_t nbytes = (uint32_t)((size_t)w * h * 4);
    const uint32_t ifd_off = 8 + nbytes;
    uint8_t hdr[8] = {'I', 'I', 42, 0, 0, 0, 0, 0};
    std::memcpy(hdr + 4, &ifd_off, 4);
    std::fwrite(hdr, 1, 8, f);
    std::fwrite(data, 4, (size_t)w * h, f);
    struct Entry { uint16_t tag, type; uint32_t count, value; };
    const Entry e[] = {
        {256, 4, 1, (uint32_t)w}, {257, 4, 1, (uint32_t)h}, {258, 3, 1, 32}, {259, 3, 1, 1},
        {262, 3, 1, 1}, {273, 4, 1, 8}, {277, 3, 1, 1}, {278, 4, 1, (uint32_t)h},
        {279, 4, 1, nbytes}, {339, 3, 1, 3},   // SampleFormat = IEEE float
    };
    const uint16_t n = sizeof(e) / sizeof(e[0]);
    std::fwrite(&n, 2, 1, f);
    for (const Entry& x : e) {
        std::fwrite(&x.tag, 2, 1, f);
        std::fwrite(&x.type, 2, 1, f);
        std::fwrite(&x.count, 4, 1, f);
        std::fwrite(&x.value, 4, 1, f);
    }
    const uint32_t next = 0;
    std::fwrite(&next, 4, 1, f);
    const bool ok = std::ferror(f) == 0;
    std::fclose(f);
    return ok;
}

}  // namespace imio
