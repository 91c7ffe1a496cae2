// Image file I/O of the host: Radiance RGBE write/read with the reference's exact encoding
// (core/image.cc:62-88,270-435) and an 8-bit PNG writer for `ldrfilm` (stored-deflate, no zlib).
#include <algorithm>
#include <cstring>
#include <fstream>

#include "core.h"

namespace spica {

void Image::saveHdr(const std::string& file) const {
    // HDRPixel (core/image.cc:60-88)
    std::vector<unsigned char> rgbe((size_t)width * height * 4);
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            const double* c = pixel(x, y);
            unsigned char* q = &rgbe[((size_t)y * width + x) * 4];
            double d = std::max(c[0], std::max(c[1], c[2]));
            if (!(d > 1.0e-32)) { q[0] = q[1] = q[2] = q[3] = 0; continue; }      // image.cc:72-76
            int ie;
            const double m = frexp(d, &ie);
            d = m * 256.0 / d;
            q[0] = (unsigned char)(c[0] * d); q[1] = (unsigned char)(c[1] * d); q[2] = (unsigned char)(c[2] * d);
            q[3] = (unsigned char)(ie + 128);
        }
    }
    writeHdrRgbe(file, width, height, rgbe.data());
}

void Image::writeHdrRgbe(const std::string& file, int width, int height, const unsigned char* rgbe) {
    std::ofstream ofs(file, std::ios::out | std::ios::binary);
    if (!ofs.is_open()) FatalError("Failed to open file: %s", file.c_str());
    char buf[256];
    // header text as the reference writes it (core/image.cc:396-407)
    snprintf(buf, sizeof(buf), "#?RADIANCE\n# Made with 100%% pure HDR Shop\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0000000000000\n\n-Y %d +X %d\n", height, width);
    ofs.write(buf, strlen(buf));
    std::vector<unsigned char> out;
    for (int y = 0; y < height; y++) {
        const unsigned char* line = rgbe + (size_t)y * width * 4;
        out.push_back(0x02); out.push_back(0x02); out.push_back((width >> 8) & 0xff); out.push_back(width & 0xff);
        for (int c = 0; c < 4; c++) {
            for (int cur = 0; cur < width;) {
                const int mv = std::min(127, width - cur);
                out.push_back((unsigned char)mv);
                for (int j = cur; j < cur + mv; j++) out.push_back(line[(size_t)j * 4 + c]);
                cur += mv;
            }
        }
    }
    ofs.write((const char*)out.data(), (std::streamsize)out.size());
}

Image Image::loadHdr(const std::string& file) {
    std::ifstream ifs(file, std::ios::in | std::ios::binary);
    if (!ifs.is_open()) FatalError("Failed to open file: %s", file.c_str());
    std::string line;
    std::getline(ifs, line);
    if (line.compare(0, 2, "#?") != 0) FatalError("Invalid HDR file: %s", file.c_str());
    bool rle = true;
    while (std::getline(ifs, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) break;
        if (line.find("FORMAT=") == 0 && line.find("rle_rgbe") == std::string::npos) FatalError("Unsupported HDR format: %s", line.c_str());
    }
    std::getline(ifs, line);
    char by[17], bx[17]; int w = 0, h = 0;
    if (sscanf(line.c_str(), "%16s %d %16s %d", by, &h, bx, &w) != 4 || strcmp(by, "-Y") != 0 || strcmp(bx, "+X") != 0)
        FatalError("Failed to parse HDRI size: %s", line.c_str());
    std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
    std::vector<unsigned char> tmp((size_t)w * h * 4, 0);
    size_t idx = 0;
    (void)rle;
    for (int y = 0; y < h; y++) {
        if (idx + 4 > bytes.size()) break;
        if (bytes[idx] == 2 && bytes[idx + 1] == 2 && (((int)bytes[idx + 2] << 8) | bytes[idx + 3]) == w) {
            idx += 4;                                       // new-style RLE scanline (image.cc:327-366)
            for (int c = 0; c < 4; c++) {
                for (int x = 0; x < w && idx < bytes.size();) {
                    const int info = bytes[idx++];
                    if (info <= 128) { for (int i = 0; i < info && x < w && idx < bytes.size(); i++) tmp[((size_t)y * w + x++) * 4 + c] = bytes[idx++]; }
                    else { const unsigned char d = bytes[idx++]; for (int i = 0; i < info - 128 && x < w; i++) tmp[((size_t)y * w + x++) * 4 + c] = d; }
                }
            }
        } else {                                            // flat RGBE
            for (int x = 0; x < w && idx + 4 <= bytes.size(); x++, idx += 4) memcpy(&tmp[((size_t)y * w + x) * 4], &bytes[idx], 4);
        }
    }
    Image img(w, h);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const int e = tmp[i * 4 + 3];
        const double s = std::pow(2.0, e - 128.0) / 256.0;                          // image.cc:373-376
        img.rgb[i * 3] = tmp[i * 4] * s; img.rgb[i * 3 + 1] = tmp[i * 4 + 1] * s; img.rgb[i * 3 + 2] = tmp[i * 4 + 2] * s;
    }
    return img;
}

Image Image::fromFile(const std::string& file) {
    const size_t p = file.find_last_of('.');
    const std::string ext = p == std::string::npos ? "" : file.substr(p);
    if (ext == ".hdr") return loadHdr(file);
    FatalError("Unsupported image type for this host (only .hdr environment maps): %s", file.c_str());
}

namespace {
uint32_t crc32(const unsigned char* d, size_t n, uint32_t crc = 0) {
    static uint32_t table[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ d[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}
void be32(std::vector<unsigned char>& v, uint32_t x) { v.push_back(x >> 24); v.push_back((x >> 16) & 0xff); v.push_back((x >> 8) & 0xff); v.push_back(x & 0xff); }
void chunk(std::ofstream& ofs, const char* type, const std::vector<unsigned char>& data) {
    std::vector<unsigned char> b;
    be32(b, (uint32_t)data.size());
    std::vector<unsigned char> td(type, type + 4);
    td.insert(td.end(), data.begin(), data.end());
    b.insert(b.end(), td.begin(), td.end());
    be32(b, crc32(td.data(), td.size()));
    ofs.write((const char*)b.data(), (std::streamsize)b.size());
}
}  // namespace

void Image::savePng(const std::string& file) const {
    std::vector<unsigned char> rgb((size_t)width * height * 3);
    for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) for (int c = 0; c < 3; c++) {
        const double v = std::min(1.0, std::max(0.0, pixel(x, y)[c]));                                 // Image::toByte (core/image.cc:484-487)
        rgb[((size_t)y * width + x) * 3 + c] = (unsigned char)(v * 255.0);
    }
    writePngRgb8(file, width, height, rgb.data());
}

void Image::writePngRgb8(const std::string& file, int width, int height, const unsigned char* rgb) {
    std::ofstream ofs(file, std::ios::out | std::ios::binary);
    if (!ofs.is_open()) FatalError("Failed to open file: %s", file.c_str());
    const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    ofs.write((const char*)sig, 8);
    std::vector<unsigned char> ihdr;
    be32(ihdr, (uint32_t)width); be32(ihdr, (uint32_t)height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(ofs, "IHDR", ihdr);
    std::vector<unsigned char> raw;
    raw.reserve((size_t)height * (width * 3 + 1));
    for (int y = 0; y < height; y++) {
        raw.push_back(0);
        raw.insert(raw.end(), rgb + (size_t)y * width * 3, rgb + (size_t)(y + 1) * width * 3);
    }
    std::vector<unsigned char> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (unsigned char c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size() || off == 0;) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        const bool last = off + n >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back(n & 0xff); z.push_back((n >> 8) & 0xff); z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
        z.insert(z.end(), raw.begin() + (long)off, raw.begin() + (long)(off + n));
        off += n;
        if (last) break;
    }
    be32(z, (b << 16) | a);
    chunk(ofs, "IDAT", z);
    chunk(ofs, "IEND", {});
}

}  // namespace spica
