// Data front ends of the reference's examples, host side only:
//   * gzip + IDX (examples/fashion_mnist/main.rs:13-72: load_gz_bytes, read_images_info / read_labels_info,
//     unpack_images = byte / 255 per selected image, unpack_labels = label byte as f32);
//   * baseline JPEG -> RGB8 (examples/image_fit/main.rs:278-282 loads data/images/cat.jpg, a 512 x 512 baseline 4:2:2 file,
//     through stb_image, a third-party dependency absent from the reference tree): sequential Huffman decoding, the IJG
//     "islow" integer inverse DCT, triangle-filter chroma upsampling for 2:1 ratios, the IJG fixed-point YCbCr -> RGB
//     conversion; checked against libjpeg (Pillow) in tests/test_cpu_host_io.py.  stb_image's own rounding may differ from
//     libjpeg's in the last bit of a pixel: parity with stb_image is UNPINNED;
//   * PPM output of a predicted image (the reference writes a JPEG through stb_image_write, main.rs:421-435; only the
//     float -> byte rule (x * 255 + 0.5 clamped) is part of the numerical path) and CSV statistics rows.
#include "host_io.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "shape.hpp"  // DSC_CHECK

namespace descent {

std::vector<uint8_t> read_file_bytes(const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    DSC_CHECK(f != nullptr, "cannot open '" << path << "'");
    std::vector<uint8_t> bytes;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + n);
    std::fclose(f);
    return bytes;
}

std::vector<uint8_t> gunzip(const uint8_t* data, size_t size) {
    z_stream s;
    std::memset(&s, 0, sizeof(s));
    DSC_CHECK(inflateInit2(&s, 16 + MAX_WBITS) == Z_OK, "inflateInit2 failed");
    s.next_in = const_cast<Bytef*>(data);
    s.avail_in = (uInt)size;
    std::vector<uint8_t> out;
    uint8_t buf[1 << 16];
    int rc = Z_OK;
    while (rc != Z_STREAM_END) {
        s.next_out = buf;
        s.avail_out = sizeof(buf);
        rc = inflate(&s, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) {
            inflateEnd(&s);
            fail("gzip stream is corrupt");
        }
        out.insert(out.end(), buf, buf + (sizeof(buf) - s.avail_out));
        if (rc == Z_OK && s.avail_in == 0 && s.avail_out != 0) {
            inflateEnd(&s);
            fail("gzip stream is truncated");
        }
    }
    inflateEnd(&s);
    return out;
}

std::vector<uint8_t> load_gz_bytes(const std::string& path) {
    auto raw = read_file_bytes(path);
    return gunzip(raw.data(), raw.size());
}

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

IdxImagesInfo read_images_info(const uint8_t* bytes, size_t size) {
    DSC_CHECK(size >= 16 && be32(bytes) == 2051, "not an IDX image file (magic 2051)");
    IdxImagesInfo info{be32(bytes + 4), be32(bytes + 8), be32(bytes + 12), 16};
    DSC_CHECK(size >= 16 + (size_t)info.images * info.rows * info.cols, "IDX image file is shorter than its header says");
    return info;
}

IdxLabelsInfo read_labels_info(const uint8_t* bytes, size_t size) {
    DSC_CHECK(size >= 8 && be32(bytes) == 2049, "not an IDX label file (magic 2049)");
    IdxLabelsInfo info{be32(bytes + 4), 8};
    DSC_CHECK(size >= 8 + (size_t)info.items, "IDX label file is shorter than its header says");
    return info;
}

void unpack_images(const uint8_t* bytes, size_t size, const uint64_t* indices, size_t count, float* out) {
    const IdxImagesInfo info = read_images_info(bytes, size);
    const size_t pixels = (size_t)info.rows * info.cols;
    for (size_t i = 0; i < count; ++i) {
        DSC_CHECK(indices[i] < info.images, "image index " << indices[i] << " out of range");
        const uint8_t* src = bytes + info.data_offset + indices[i] * pixels;
        for (size_t p = 0; p < pixels; ++p) out[i * pixels + p] = (float)src[p] / 255.0f;
    }
}

void unpack_labels(const uint8_t* bytes, size_t size, const uint64_t* indices, size_t count, float* out) {
    const IdxLabelsInfo info = read_labels_info(bytes, size);
    for (size_t i = 0; i < count; ++i) {
        DSC_CHECK(indices[i] < info.items, "label index " << indices[i] << " out of range");
        out[i] = (float)bytes[info.data_offset + indices[i]];
    }
}

// ---- baseline JPEG ------------------------------------------------------------------------------------------------------
namespace {

const uint8_t kZigZag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huffman {
    bool present = false;
    int mincode[17], maxcode[18], valptr[17];
    uint8_t symbols[256];
    void build(const uint8_t counts[16], const uint8_t* syms, int total) {
        std::memcpy(symbols, syms, (size_t)total);
        int code = 0, k = 0;
        for (int len = 1; len <= 16; ++len) {
            valptr[len] = k;
            mincode[len] = code;
            code += counts[len - 1];
            k += counts[len - 1];
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t bits = 0;
    int count = 0;
    bool hit_marker = false;
    int bit() {
        if (count == 0) {
            uint8_t b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xff) {
                    if (p < end && *p == 0x00) ++p;            // stuffed zero
                    else { hit_marker = true; --p; b = 0; }  // a marker: feed zeros until the caller handles it
                }
            }
            bits = b;
            count = 8;
        }
        --count;
        return (bits >> count) & 1;
    }
    int receive(int n) {
        int v = 0;
        for (int i = 0; i < n; ++i) v = (v << 1) | bit();
        return v;
    }
    void reset() { bits = 0; count = 0; hit_marker = false; }
};

int decode_symbol(BitReader& br, const Huffman& h) {
    int code = 0;
    for (int len = 1; len <= 16; ++len) {
        code = (code << 1) | br.bit();
        if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) return h.symbols[h.valptr[len] + code - h.mincode[len]];
    }
    fail("corrupt JPEG: bad Huffman code");
}

int extend(int v, int n) { return n == 0 ? 0 : (v < (1 << (n - 1)) ? v - (1 << n) + 1 : v); }

// IJG jidctint.c ("islow"): CONST_BITS = 13, PASS1_BITS = 2
void idct_islow(const int* in, uint8_t* out, int stride) {
    constexpr int CB = 13, P1 = 2;
    constexpr int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137, F1961 = 16069, F2053 = 16819,
                  F2562 = 20995, F3072 = 25172;
    int ws[64];
    auto descale = [](long long x, int n) { return (int)((x + ((long long)1 << (n - 1))) >> n); };
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < 8; ++i) {
            const int* src = pass == 0 ? in + i : ws + 8 * i;
            const int step = pass == 0 ? 8 : 1;
            long long z2 = src[2 * step], z3 = src[6 * step];
            long long z1 = (z2 + z3) * F0541;
            long long tmp2 = z1 + z3 * (-F1847), tmp3 = z1 + z2 * F0765;
            z2 = src[0];
            z3 = src[4 * step];
            long long tmp0 = (z2 + z3) << CB, tmp1 = (z2 - z3) << CB;
            const long long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = src[7 * step]; tmp1 = src[5 * step]; tmp2 = src[3 * step]; tmp3 = src[1 * step];
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            long long z4 = tmp1 + tmp3;
            const long long z5 = (z3 + z4) * F1175;
            tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
            z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            const long long r[8] = {tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3};
            if (pass == 0) {
                for (int k = 0; k < 8; ++k) ws[8 * k + i] = descale(r[k], CB - P1);
            } else {
                for (int k = 0; k < 8; ++k) out[i * stride + k] = (uint8_t)std::min(255, std::max(0, descale(r[k], CB + P1 + 3) + 128));
            }
        }
    }
}

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int width = 0, height = 0;        // sample dimensions
    int stride = 0, rows = 0;         // padded to whole MCUs
    int pred = 0;
    std::vector<uint8_t> plane;
};

}  // namespace

JpegImage decode_jpeg_rgb(const uint8_t* data, size_t size) {
    DSC_CHECK(size >= 4 && data[0] == 0xff && data[1] == 0xd8, "not a JPEG file");
    uint16_t quant[4][64] = {};
    Huffman dc[4], ac[4];
    std::vector<Component> comps;
    int width = 0, height = 0, restart = 0;
    size_t pos = 2;
    auto u16 = [&](size_t at) { return (int)((data[at] << 8) | data[at + 1]); };
    for (;;) {
        DSC_CHECK(pos + 4 <= size && data[pos] == 0xff, "corrupt JPEG: marker expected");
        const int marker = data[pos + 1];
        if (marker == 0xff) { pos += 1; continue; }
        const int len = u16(pos + 2);
        DSC_CHECK(pos + 2 + len <= size, "corrupt JPEG: segment overruns the file");
        const uint8_t* seg = data + pos + 4;
        const int body = len - 2;
        if (marker == 0xdb) {  // DQT
            for (int at = 0; at < body;) {
                const int pq = seg[at] >> 4, tq = seg[at] & 15;
                DSC_CHECK(tq < 4, "corrupt JPEG: quantisation table id");
                ++at;
                for (int i = 0; i < 64; ++i) {
                    quant[tq][kZigZag[i]] = pq ? (uint16_t)((seg[at] << 8) | seg[at + 1]) : seg[at];
                    at += pq ? 2 : 1;
                }
            }
        } else if (marker == 0xc4) {  // DHT
            for (int at = 0; at < body;) {
                const int tc = seg[at] >> 4, th = seg[at] & 15;
                DSC_CHECK(tc < 2 && th < 4, "corrupt JPEG: Huffman table id");
                int total = 0;
                for (int i = 0; i < 16; ++i) total += seg[at + 1 + i];
                DSC_CHECK(total <= 256 && at + 17 + total <= body, "corrupt JPEG: Huffman table size");
                (tc ? ac[th] : dc[th]).build(seg + at + 1, seg + at + 17, total);
                at += 17 + total;
            }
        } else if (marker == 0xc0 || marker == 0xc1) {  // SOF0 / SOF1: sequential, Huffman
            DSC_CHECK(seg[0] == 8, "only 8-bit JPEG samples are supported");
            height = u16(pos + 5);
            width = u16(pos + 7);
            const int n = seg[5];
            DSC_CHECK((n == 1 || n == 3) && width > 0 && height > 0, "only greyscale and YCbCr JPEG files are supported");
            comps.resize((size_t)n);
            for (int i = 0; i < n; ++i) {
                comps[i].id = seg[6 + 3 * i];
                comps[i].h = seg[7 + 3 * i] >> 4;
                comps[i].v = seg[7 + 3 * i] & 15;
                comps[i].tq = seg[8 + 3 * i];
                DSC_CHECK(comps[i].h >= 1 && comps[i].h <= 4 && comps[i].v >= 1 && comps[i].v <= 4 && comps[i].tq < 4, "corrupt JPEG: component header");
            }
        } else if (marker == 0xc2 || (marker >= 0xc3 && marker <= 0xcf && marker != 0xc4 && marker != 0xc8 && marker != 0xcc)) {
            fail("progressive / lossless / arithmetic-coded JPEG files are not supported (baseline only)");
        } else if (marker == 0xdd) {
            restart = u16(pos + 4);
        } else if (marker == 0xda) {  // SOS
            DSC_CHECK(!comps.empty(), "corrupt JPEG: scan before frame header");
            const int ns = seg[0];
            DSC_CHECK(ns == (int)comps.size(), "only single-scan (interleaved) baseline JPEG files are supported");
            for (int i = 0; i < ns; ++i) {
                const int cid = seg[1 + 2 * i];
                bool found = false;
                for (auto& c : comps)
                    if (c.id == cid) { c.td = seg[2 + 2 * i] >> 4; c.ta = seg[2 + 2 * i] & 15; found = true; }
                DSC_CHECK(found, "corrupt JPEG: scan names an unknown component");
            }
            pos += 2 + (size_t)len;
            break;
        } else if (marker == 0xd9) {
            fail("corrupt JPEG: no scan");
        }
        pos += 2 + (size_t)len;
    }
    int hmax = 1, vmax = 1;
    for (const auto& c : comps) { hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v); }
    const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
    const int mcus_x = (width + mcu_w - 1) / mcu_w, mcus_y = (height + mcu_h - 1) / mcu_h;
    for (auto& c : comps) {
        c.width = (width * c.h + hmax - 1) / hmax;
        c.height = (height * c.v + vmax - 1) / vmax;
        c.stride = mcus_x * c.h * 8;
        c.rows = mcus_y * c.v * 8;
        c.plane.assign((size_t)c.stride * c.rows, 0);
        DSC_CHECK(dc[c.td].present && ac[c.ta].present, "corrupt JPEG: scan uses an undefined Huffman table");
    }
    BitReader br{data + pos, data + size};
    int until_restart = restart;
    for (int my = 0; my < mcus_y; ++my) {
        for (int mx = 0; mx < mcus_x; ++mx) {
            if (restart && until_restart == 0) {
                // byte-align, expect RSTn
                br.reset();
                while (br.p + 1 < br.end && !(br.p[0] == 0xff && br.p[1] >= 0xd0 && br.p[1] <= 0xd7)) ++br.p;
                DSC_CHECK(br.p + 1 < br.end, "corrupt JPEG: restart marker missing");
                br.p += 2;
                for (auto& c : comps) c.pred = 0;
                until_restart = restart;
            }
            for (auto& c : comps) {
                for (int by = 0; by < c.v; ++by) {
                    for (int bx = 0; bx < c.h; ++bx) {
                        int coef[64] = {};
                        const int t = decode_symbol(br, dc[c.td]);
                        DSC_CHECK(t <= 11, "corrupt JPEG: DC category");
                        c.pred += extend(br.receive(t), t);
                        coef[0] = c.pred * quant[c.tq][0];
                        for (int k = 1; k < 64;) {
                            const int rs = decode_symbol(br, ac[c.ta]);
                            const int r = rs >> 4, s = rs & 15;
                            if (s == 0) {
                                if (r != 15) break;
                                k += 16;
                                continue;
                            }
                            k += r;
                            DSC_CHECK(k < 64, "corrupt JPEG: AC run past the block");
                            coef[kZigZag[k]] = extend(br.receive(s), s) * quant[c.tq][kZigZag[k]];
                            ++k;
                        }
                        idct_islow(coef, c.plane.data() + (size_t)((my * c.v + by) * 8) * c.stride + (mx * c.h + bx) * 8, c.stride);
                    }
                }
            }
            if (restart) --until_restart;
        }
    }
    // chroma upsampling to the luma grid.  2:1 ratios use the IJG "fancy" triangle filters (jdsample.c h2v1 / h2v2); other
    // ratios replicate samples.
    auto upsample = [&](const Component& c) {
        std::vector<uint8_t> full((size_t)width * height);
        const int fx = hmax / c.h, fy = vmax / c.v;
        auto at = [&](int y, int x) { return (int)c.plane[(size_t)std::min(std::max(y, 0), c.height - 1) * c.stride + std::min(std::max(x, 0), c.width - 1)]; };
        if (fx == 1 && fy == 1) {
            for (int y = 0; y < height; ++y)
                for (int x = 0; x < width; ++x) full[(size_t)y * width + x] = (uint8_t)at(y, x);
        } else if (fx == 2 && fy == 1 && hmax % c.h == 0) {
            for (int y = 0; y < height; ++y)
                for (int x = 0; x < width; ++x) {
                    const int i = x >> 1;
                    int v;
                    if (c.width == 1) v = at(y, 0);
                    else if ((x & 1) == 0) v = i == 0 ? at(y, 0) : (at(y, i) * 3 + at(y, i - 1) + 1) >> 2;
                    else v = i == c.width - 1 ? at(y, i) : (at(y, i) * 3 + at(y, i + 1) + 2) >> 2;
                    full[(size_t)y * width + x] = (uint8_t)v;
                }
        } else if (fx == 2 && fy == 2 && hmax % c.h == 0 && vmax % c.v == 0) {
            for (int y = 0; y < height; ++y) {
                const int j = y >> 1, other = (y & 1) ? j + 1 : j - 1;  // nearer row weighs 3, the other neighbour 1
                auto colsum = [&](int i) { return at(j, i) * 3 + at(other, i); };
                for (int x = 0; x < width; ++x) {
                    const int i = x >> 1;
                    int v;
                    if (c.width == 1) v = (colsum(0) * 4 + 8) >> 4;
                    else if ((x & 1) == 0) v = i == 0 ? (colsum(0) * 4 + 8) >> 4 : (colsum(i) * 3 + colsum(i - 1) + 8) >> 4;
                    else v = i == c.width - 1 ? (colsum(i) * 4 + 7) >> 4 : (colsum(i) * 3 + colsum(i + 1) + 7) >> 4;
                    full[(size_t)y * width + x] = (uint8_t)v;
                }
            }
        } else {
            for (int y = 0; y < height; ++y)
                for (int x = 0; x < width; ++x) full[(size_t)y * width + x] = (uint8_t)at(y * c.v / vmax, x * c.h / hmax);
        }
        return full;
    };
    JpegImage image;
    image.width = width;
    image.height = height;
    image.rgb.resize((size_t)width * height * 3);
    const std::vector<uint8_t> yp = upsample(comps[0]);
    if (comps.size() == 1) {
        for (size_t i = 0; i < yp.size(); ++i) image.rgb[3 * i] = image.rgb[3 * i + 1] = image.rgb[3 * i + 2] = yp[i];
        return image;
    }
    const std::vector<uint8_t> cb = upsample(comps[1]), cr = upsample(comps[2]);
    // IJG jdcolor.c: 16-bit fixed point tables
    auto fix = [](double x) { return (long)(x * 65536.0 + 0.5); };
    auto clamp = [](long v) { return (uint8_t)std::min<long>(255, std::max<long>(0, v)); };
    for (size_t i = 0; i < yp.size(); ++i) {
        const long y = yp[i], b = (long)cb[i] - 128, r = (long)cr[i] - 128;
        image.rgb[3 * i] = clamp(y + ((fix(1.40200) * r + 32768) >> 16));
        image.rgb[3 * i + 1] = clamp(y + ((-fix(0.34414) * b + 32768 - fix(0.71414) * r) >> 16));
        image.rgb[3 * i + 2] = clamp(y + ((fix(1.77200) * b + 32768) >> 16));
    }
    return image;
}

void write_ppm(const std::string& path, const float* rgb, int width, int height) {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    DSC_CHECK(f != nullptr, "cannot create '" << path << "'");
    std::fprintf(f, "P6\n%d %d\n255\n", width, height);
    std::vector<uint8_t> row((size_t)width * 3);
    for (int y = 0; y < height; ++y) {
        for (int i = 0; i < width * 3; ++i) {
            const float v = rgb[(size_t)y * width * 3 + i] * 255.0f + 0.5f;  // examples/image_fit/main.rs:423-426
            row[(size_t)i] = (uint8_t)std::min(255.0f, std::max(0.0f, v));
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
}

}  // namespace descent
