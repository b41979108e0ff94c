// ASan + UBSan over the PNG codec (portal_b200/csrc/host/ph_png.cpp): the decoder reads files from disk, so it must survive
// anything.  Encodes synthetic frames, then decodes (a) the files as they are -- round trip must be exact --, (b) every
// prefix, (c) thousands of byte mutations with the chunk CRCs repaired so that the damage reaches the zlib / filter /
// palette layers, (d) raw zlib streams with mutations.  Also decodes any PNG files given on the command line with the same
// mutations.  Built and run by tools/sanitize_host.sh; prints counts.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "ph_png.h"

static void fix_crcs(std::vector<uint8_t>& f) {
    size_t pos = 8;
    while (pos + 12 <= f.size()) {
        const uint32_t n = (uint32_t(f[pos]) << 24) | (uint32_t(f[pos + 1]) << 16) | (uint32_t(f[pos + 2]) << 8) | f[pos + 3];
        if (n > f.size() || pos + 12 + n > f.size()) return;
        const uint32_t c = ph::crc32(f.data() + pos + 4, n + 4);
        for (int k = 0; k < 4; k++) f[pos + 8 + n + k] = uint8_t(c >> (24 - 8 * k));
        pos += 12 + n;
    }
}

int main(int argc, char** argv) {
    std::mt19937 rng(777);
    long ok = 0, refused = 0;
    std::vector<std::vector<uint8_t>> files;
    for (int t = 0; t < 6; t++) {
        const int w = 1 + int(rng() % 70), h = 1 + int(rng() % 40);
        std::vector<uint8_t> px(size_t(w) * h * 4);
        for (size_t i = 0; i < px.size(); i++) px[i] = t % 2 ? uint8_t(rng()) : uint8_t((i / 4 % w) * 3 + (i % 4) * 40);
        std::vector<uint8_t> f;
        ph::png_encode_rgba8(px.data(), w, h, f);
        std::vector<uint8_t> back;
        int bw = 0, bh = 0;
        std::string err;
        if (!ph::png_decode(f.data(), f.size(), back, bw, bh, err) || bw != w || bh != h || back != px) { std::printf("ROUND TRIP FAILED: %s\n", err.c_str()); return 1; }
        files.push_back(f);
    }
    for (int i = 1; i < argc; i++) {
        std::ifstream in(argv[i], std::ios::binary);
        std::ostringstream ss;
        ss << in.rdbuf();
        const std::string s = ss.str();
        if (s.size() > (4u << 20)) continue;                   // keep the run short: small files only
        files.emplace_back(s.begin(), s.end());
    }
    for (const auto& f : files) {
        std::vector<uint8_t> out;
        int w = 0, h = 0;
        std::string err;
        for (size_t cut = 0; cut < f.size(); cut += 1 + f.size() / 400) (ph::png_decode(f.data(), cut, out, w, h, err) ? ok : refused)++;
        for (int m = 0; m < 1500; m++) {
            std::vector<uint8_t> g = f;
            const int edits = 1 + int(rng() % 3);
            for (int e = 0; e < edits; e++) g[rng() % g.size()] ^= uint8_t(1 + rng() % 255);
            if (m % 3) fix_crcs(g);
            (ph::png_decode(g.data(), g.size(), out, w, h, err) ? ok : refused)++;
        }
    }
    for (int t = 0; t < 40; t++) {                               // the zlib layer on its own
        std::vector<uint8_t> raw(size_t(rng() % 5000)), z, back;
        for (auto& b : raw) b = t % 2 ? uint8_t(rng()) : uint8_t(rng() % 4);
        ph::zlib_deflate(raw.data(), raw.size(), z);
        std::string err;
        if (!ph::zlib_inflate(z.data(), z.size(), back, err) || back != raw) { std::printf("ZLIB ROUND TRIP FAILED: %s\n", err.c_str()); return 1; }
        for (int m = 0; m < 200; m++) {
            std::vector<uint8_t> g = z;
            g[rng() % g.size()] ^= uint8_t(1 + rng() % 255);
            (ph::zlib_inflate(g.data(), g.size(), back, err) ? ok : refused)++;
        }
    }
    std::printf("png fuzz: %ld decoded, %ld refused, no crash\n", ok, refused);
    return 0;
}
