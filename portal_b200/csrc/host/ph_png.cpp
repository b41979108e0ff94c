// PNG codec of the host side (see ph_png.h): RFC 2083 (PNG), RFC 1950 (zlib), RFC 1951 (DEFLATE) written out.
#include "ph_png.h"

#include <algorithm>
#include <cstring>

namespace ph {

// ------------------------------------------------------------------------------------ checksums
namespace {
struct CrcTable {
    uint32_t t[256];
    CrcTable() {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            t[i] = c;
        }
    }
};
}  // namespace

uint32_t crc32(const uint8_t* d, size_t n, uint32_t crc) {
    static const CrcTable T;
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = T.t[(crc ^ d[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}

uint32_t adler32(const uint8_t* d, size_t n) {
    uint32_t a = 1, b = 0;
    while (n) {
        const size_t k = n < 5552 ? n : 5552;      // the largest run for which b cannot overflow 32 bits
        for (size_t i = 0; i < k; i++) { a += d[i]; b += a; }
        a %= 65521u;
        b %= 65521u;
        d += k;
        n -= k;
    }
    return (b << 16) | a;
}

// ------------------------------------------------------------------------------------ inflate
namespace {

struct BitReader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint64_t buf = 0;
    int cnt = 0;
    bool over = false;      // asked for bits past the end
    BitReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
    void fill() {
        while (cnt <= 56 && pos < n) { buf |= uint64_t(p[pos++]) << cnt; cnt += 8; }
    }
    uint32_t peek(int k) {  // k <= 32; missing bits read as 0 (drop() flags the overrun)
        if (cnt < k) fill();
        return uint32_t(buf & ((uint64_t(1) << k) - 1));
    }
    void drop(int k) {
        if (cnt < k) { over = true; cnt = 0; buf = 0; return; }
        buf >>= k;
        cnt -= k;
    }
    uint32_t bits(int k) {
        const uint32_t v = peek(k);
        drop(k);
        return v;
    }
    void align() { drop(cnt & 7); }
};

constexpr int kFast = 10;

struct Huffman {
    uint16_t count[16] = {0};
    uint16_t symbol[320] = {0};
    uint16_t fast[1 << kFast];   // (symbol << 4) | length, 0 = longer than kFast bits (or unused)
    // canonical code from code lengths (RFC 1951 3.2.2); false = over-subscribed
    bool build(const uint8_t* lengths, int n) {
        std::memset(count, 0, sizeof count);
        std::memset(fast, 0, sizeof fast);
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        count[0] = 0;
        int left = 1;
        for (int len = 1; len <= 15; len++) {
            left <<= 1;
            left -= count[len];
            if (left < 0) return false;
        }
        uint16_t offs[16];
        offs[1] = 0;
        for (int len = 1; len < 15; len++) offs[len + 1] = uint16_t(offs[len] + count[len]);
        for (int i = 0; i < n; i++)
            if (lengths[i]) symbol[offs[lengths[i]]++] = uint16_t(i);
        uint32_t next[16] = {0}, code = 0;
        for (int len = 1; len <= 15; len++) {
            code = (code + count[len - 1]) << 1;
            next[len] = code;
        }
        for (int i = 0; i < n; i++) {
            const int len = lengths[i];
            if (!len) continue;
            const uint32_t c = next[len]++;
            if (len > kFast) continue;
            uint32_t rev = 0;
            for (int b = 0; b < len; b++) rev |= ((c >> b) & 1u) << (len - 1 - b);   // the stream carries codes MSB first
            for (uint32_t k = rev; k < (1u << kFast); k += 1u << len) fast[k] = uint16_t((i << 4) | len);
        }
        return true;
    }
    int decode(BitReader& br) const {
        const uint16_t e = fast[br.peek(kFast)];
        if (e) {
            br.drop(e & 15);
            return e >> 4;
        }
        int code = 0, first = 0, index = 0;           // walk the canonical code one bit at a time
        for (int len = 1; len <= 15; len++) {
            code |= int(br.bits(1));
            const int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        return -1;
    }
};

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

struct FixedTables {
    Huffman lit, dist;
    FixedTables() {
        uint8_t l[288];
        for (int i = 0; i < 288; i++) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        lit.build(l, 288);
        uint8_t d[30];
        for (int i = 0; i < 30; i++) d[i] = 5;
        dist.build(d, 30);
    }
};

bool inflate_block(BitReader& br, const Huffman& lit, const Huffman& dist, std::vector<uint8_t>& out, std::string& err) {
    for (;;) {
        const int sym = lit.decode(br);
        if (br.over) { err = "deflate stream ends inside a block"; return false; }
        if (sym < 0 || sym > 285) { err = "bad literal/length code"; return false; }
        if (sym < 256) { out.push_back(uint8_t(sym)); continue; }
        if (sym == 256) return true;
        const int li = sym - 257;
        const size_t len = kLenBase[li] + br.bits(kLenExtra[li]);
        const int ds = dist.decode(br);
        if (ds < 0 || ds > 29) { err = "bad distance code"; return false; }
        const size_t d = kDistBase[ds] + br.bits(kDistExtra[ds]);
        if (br.over) { err = "deflate stream ends inside a block"; return false; }
        if (d > out.size()) { err = "distance reaches before the start of the data"; return false; }
        const size_t at = out.size();
        out.resize(at + len);
        uint8_t* o = out.data() + at;
        const uint8_t* s = o - d;
        for (size_t i = 0; i < len; i++) o[i] = s[i];      // may overlap: byte by byte, forwards
    }
}

}  // namespace

bool zlib_inflate(const uint8_t* data, size_t len, std::vector<uint8_t>& out, std::string& err) {
    out.clear();
    if (len < 6) { err = "zlib stream too short"; return false; }
    const unsigned cmf = data[0], flg = data[1];
    if ((cmf & 15u) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31u != 0) { err = "not a zlib (deflate) stream"; return false; }
    if (flg & 0x20u) { err = "zlib preset dictionary is not allowed in PNG"; return false; }
    static const FixedTables fixed;
    BitReader br(data + 2, len - 2);
    for (bool last = false; !last;) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (br.over) { err = "deflate stream ends before its last block"; return false; }
        if (type == 0) {
            br.align();
            const uint32_t n = br.bits(16), nn = br.bits(16);
            if (br.over || n != ((~nn) & 0xFFFFu)) { err = "bad stored block"; return false; }
            const size_t at = out.size();
            out.resize(at + n);
            for (uint32_t i = 0; i < n; i++) out[at + i] = uint8_t(br.bits(8));
            if (br.over) { err = "stored block runs past the end"; return false; }
        } else if (type == 1) {
            if (!inflate_block(br, fixed.lit, fixed.dist, out, err)) return false;
        } else if (type == 2) {
            const int hlit = int(br.bits(5)) + 257, hdist = int(br.bits(5)) + 1, hclen = int(br.bits(4)) + 4;
            if (hlit > 286 || hdist > 30) { err = "bad code counts"; return false; }
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; i++) cl[order[i]] = uint8_t(br.bits(3));
            Huffman clh;
            if (!clh.build(cl, 19)) { err = "bad code-length code"; return false; }
            uint8_t lengths[320] = {0};
            for (int i = 0; i < hlit + hdist;) {
                const int s = clh.decode(br);
                if (br.over || s < 0 || s > 18) { err = "bad code lengths"; return false; }
                if (s < 16) { lengths[i++] = uint8_t(s); continue; }
                int rep, val = 0;
                if (s == 16) {
                    if (i == 0) { err = "repeat with nothing to repeat"; return false; }
                    val = lengths[i - 1];
                    rep = 3 + int(br.bits(2));
                } else if (s == 17) rep = 3 + int(br.bits(3));
                else rep = 11 + int(br.bits(7));
                if (i + rep > hlit + hdist) { err = "code lengths overrun"; return false; }
                while (rep--) lengths[i++] = uint8_t(val);
            }
            if (lengths[256] == 0) { err = "no end-of-block code"; return false; }
            Huffman lit, dist;
            if (!lit.build(lengths, hlit) || !dist.build(lengths + hlit, hdist)) { err = "over-subscribed code"; return false; }
            if (!inflate_block(br, lit, dist, out, err)) return false;
        } else {
            err = "reserved block type";
            return false;
        }
    }
    br.align();
    uint32_t want = 0;
    for (int i = 0; i < 4; i++) want = (want << 8) | br.bits(8);
    if (br.over) { err = "zlib stream has no checksum"; return false; }
    if (want != adler32(out.data(), out.size())) { err = "zlib checksum mismatch"; return false; }
    return true;
}

// ------------------------------------------------------------------------------------ deflate
namespace {

struct BitWriter {
    std::vector<uint8_t>& out;
    uint64_t buf = 0;
    int cnt = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    void put(uint32_t v, int k) {      // LSB first
        buf |= uint64_t(v) << cnt;
        cnt += k;
        while (cnt >= 8) { out.push_back(uint8_t(buf)); buf >>= 8; cnt -= 8; }
    }
    void put_code(uint32_t code, int k) {   // Huffman codes go MSB first
        uint32_t rev = 0;
        for (int b = 0; b < k; b++) rev |= ((code >> b) & 1u) << (k - 1 - b);
        put(rev, k);
    }
    void flush() {
        if (cnt) { out.push_back(uint8_t(buf)); buf = 0; cnt = 0; }
    }
};

void put_fixed_literal(BitWriter& bw, int sym) {      // RFC 1951 3.2.6
    if (sym < 144) bw.put_code(0x30u + uint32_t(sym), 8);
    else if (sym < 256) bw.put_code(0x190u + uint32_t(sym - 144), 9);
    else if (sym < 280) bw.put_code(uint32_t(sym - 256), 7);
    else bw.put_code(0xC0u + uint32_t(sym - 280), 8);
}

}  // namespace

void zlib_deflate(const uint8_t* d, size_t n, std::vector<uint8_t>& out) {
    out.clear();
    out.reserve(n / 2 + 64);
    out.push_back(0x78);
    out.push_back(0x9C);
    BitWriter bw(out);
    bw.put(1, 1);      // last block
    bw.put(1, 2);      // fixed Huffman codes
    constexpr int kHashBits = 15, kWindow = 32768, kChain = 16;
    std::vector<int32_t> head(size_t(1) << kHashBits, -1), prev(kWindow, -1);
    auto hash = [&](size_t i) { return ((uint32_t(d[i]) << 10) ^ (uint32_t(d[i + 1]) << 5) ^ uint32_t(d[i + 2])) & ((1u << kHashBits) - 1); };
    auto insert = [&](size_t i) {
        const uint32_t h = hash(i);
        prev[i & (kWindow - 1)] = head[h];
        head[h] = int32_t(i);
    };
    size_t i = 0;
    while (i < n) {
        size_t best_len = 0, best_dist = 0;
        if (i + 3 <= n) {
            const size_t max_len = std::min<size_t>(258, n - i);
            int32_t cand = head[hash(i)];
            for (int probe = 0; probe < kChain && cand >= 0 && i - size_t(cand) <= kWindow; probe++) {
                const size_t c = size_t(cand);
                if (d[c + best_len] == d[i + best_len] || best_len == 0) {
                    size_t l = 0;
                    while (l < max_len && d[c + l] == d[i + l]) l++;
                    if (l > best_len) { best_len = l; best_dist = i - c; if (l == max_len) break; }
                }
                const int32_t nx = prev[c & (kWindow - 1)];
                if (nx >= cand) break;      // the slot was overwritten by a newer position: end of this chain
                cand = nx;
            }
        }
        if (best_len >= 3) {
            int li = 28;
            while (kLenBase[li] > best_len) li--;
            put_fixed_literal(bw, 257 + li);
            bw.put(uint32_t(best_len - kLenBase[li]), kLenExtra[li]);
            int di = 29;
            while (kDistBase[di] > best_dist) di--;
            bw.put_code(uint32_t(di), 5);
            bw.put(uint32_t(best_dist - kDistBase[di]), kDistExtra[di]);
            for (size_t k = 0; k < best_len; k++)
                if (i + k + 3 <= n) insert(i + k);
            i += best_len;
        } else {
            put_fixed_literal(bw, d[i]);
            if (i + 3 <= n) insert(i);
            i++;
        }
    }
    put_fixed_literal(bw, 256);
    bw.flush();
    const uint32_t a = adler32(d, n);
    for (int s = 24; s >= 0; s -= 8) out.push_back(uint8_t(a >> s));
}

// ------------------------------------------------------------------------------------ PNG
namespace {

const uint8_t kSignature[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};

uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

void put_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
    for (int s = 24; s >= 0; s -= 8) out.push_back(uint8_t(uint32_t(n) >> s));
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    const uint32_t c = crc32(out.data() + at, n + 4);
    for (int s = 24; s >= 0; s -= 8) out.push_back(uint8_t(c >> s));
}

}  // namespace

bool png_decode(const uint8_t* data, size_t len, std::vector<uint8_t>& rgba, int& width, int& height, std::string& err) {
    rgba.clear();
    width = height = 0;
    if (len < 8 || std::memcmp(data, kSignature, 8) != 0) { err = "not a PNG file (bad signature)"; return false; }
    size_t pos = 8;
    bool have_ihdr = false, have_end = false;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0;
    std::vector<uint8_t> idat, plte, trns;
    while (pos + 12 <= len && !have_end) {
        const uint32_t n = be32(data + pos);
        const uint8_t* type = data + pos + 4;
        if (n > 0x7FFFFFFFu || pos + 12 + size_t(n) > len) { err = "chunk runs past the end of the file"; return false; }
        const uint8_t* body = data + pos + 8;
        if (crc32(type, size_t(n) + 4) != be32(body + n)) { err = std::string("CRC mismatch in chunk ") + std::string(reinterpret_cast<const char*>(type), 4); return false; }
        auto tag = [&](const char* t) { return std::memcmp(type, t, 4) == 0; };
        if (!have_ihdr && !tag("IHDR")) { err = "first chunk is not IHDR"; return false; }
        if (tag("IHDR")) {
            if (have_ihdr || n != 13) { err = "bad IHDR"; return false; }
            have_ihdr = true;
            w = be32(body);
            h = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            if (w == 0 || h == 0 || w > 0x7FFFFFFFu || h > 0x7FFFFFFFu || uint64_t(w) * uint64_t(h) > (uint64_t(1) << 28)) { err = "unreasonable image size"; return false; }
            if (body[10] != 0 || body[11] != 0) { err = "unknown compression / filter method"; return false; }
            if (body[12] != 0) { err = "interlaced (Adam7) PNG is not supported"; return false; }
            const bool ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                            (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                            ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
            if (!ok) { err = "invalid colour type / bit depth"; return false; }
            if (depth == 16) { err = "16-bit samples are not supported"; return false; }
        } else if (tag("PLTE")) {
            if (n % 3 != 0 || n > 768) { err = "bad PLTE"; return false; }
            plte.assign(body, body + n);
        } else if (tag("tRNS")) {
            trns.assign(body, body + n);
        } else if (tag("IDAT")) {
            idat.insert(idat.end(), body, body + n);
        } else if (tag("IEND")) {
            have_end = true;
        } else if (!(type[0] & 0x20)) {
            err = std::string("unknown critical chunk ") + std::string(reinterpret_cast<const char*>(type), 4);
            return false;
        }
        pos += 12 + size_t(n);
    }
    if (!have_ihdr || !have_end) { err = "truncated PNG (no IEND)"; return false; }
    if (ctype == 3 && plte.empty()) { err = "palette image without PLTE"; return false; }
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
    const size_t row_bytes = (size_t(w) * size_t(channels) * size_t(depth) + 7) / 8;
    const size_t bpp = std::max<size_t>(1, size_t(channels) * size_t(depth) / 8);
    std::vector<uint8_t> raw;
    if (!zlib_inflate(idat.data(), idat.size(), raw, err)) return false;
    if (raw.size() != size_t(h) * (row_bytes + 1)) { err = "image data has the wrong size"; return false; }
    // ---- undo the row filters in place (RFC 2083 section 6)
    std::vector<uint8_t> zero(row_bytes, 0);
    for (size_t y = 0; y < h; y++) {
        uint8_t* row = raw.data() + y * (row_bytes + 1);
        const uint8_t ft = row[0];
        uint8_t* cur = row + 1;
        const uint8_t* up = y ? raw.data() + (y - 1) * (row_bytes + 1) + 1 : zero.data();
        switch (ft) {
            case 0: break;
            case 1: for (size_t i = bpp; i < row_bytes; i++) cur[i] = uint8_t(cur[i] + cur[i - bpp]); break;
            case 2: for (size_t i = 0; i < row_bytes; i++) cur[i] = uint8_t(cur[i] + up[i]); break;
            case 3:
                for (size_t i = 0; i < row_bytes; i++) cur[i] = uint8_t(cur[i] + ((int(i >= bpp ? cur[i - bpp] : 0) + int(up[i])) >> 1));
                break;
            case 4:
                for (size_t i = 0; i < row_bytes; i++)
                    cur[i] = uint8_t(cur[i] + paeth(i >= bpp ? cur[i - bpp] : 0, up[i], i >= bpp ? up[i - bpp] : 0));
                break;
            default: err = "unknown row filter"; return false;
        }
    }
    // ---- samples -> RGBA8
    rgba.resize(size_t(w) * size_t(h) * 4);
    const int maxv = (1 << depth) - 1;
    const int key_gray = trns.size() >= 2 ? ((trns[0] << 8) | trns[1]) : -1;
    for (size_t y = 0; y < h; y++) {
        const uint8_t* src = raw.data() + y * (row_bytes + 1) + 1;
        uint8_t* dst = rgba.data() + y * size_t(w) * 4;
        for (size_t x = 0; x < w; x++, dst += 4) {
            if (ctype == 6) {
                std::memcpy(dst, src + 4 * x, 4);
            } else if (ctype == 2) {
                dst[0] = src[3 * x]; dst[1] = src[3 * x + 1]; dst[2] = src[3 * x + 2];
                dst[3] = (trns.size() >= 6 && dst[0] == trns[1] && dst[1] == trns[3] && dst[2] == trns[5] && trns[0] == 0 && trns[2] == 0 && trns[4] == 0) ? 0 : 255;
            } else if (ctype == 4) {
                dst[0] = dst[1] = dst[2] = src[2 * x];
                dst[3] = src[2 * x + 1];
            } else {
                int v;
                if (depth == 8) v = src[x];
                else {
                    const size_t bit = x * size_t(depth);
                    v = (src[bit >> 3] >> (8 - depth - int(bit & 7))) & maxv;      // leftmost pixel in the high bits
                }
                if (ctype == 3) {
                    if (size_t(v) * 3 + 3 > plte.size()) { err = "palette index out of range"; return false; }
                    dst[0] = plte[size_t(v) * 3]; dst[1] = plte[size_t(v) * 3 + 1]; dst[2] = plte[size_t(v) * 3 + 2];
                    dst[3] = size_t(v) < trns.size() ? trns[size_t(v)] : 255;
                } else {
                    dst[0] = dst[1] = dst[2] = uint8_t(v * 255 / maxv);
                    dst[3] = v == key_gray ? 0 : 255;
                }
            }
        }
    }
    width = int(w);
    height = int(h);
    return true;
}

void png_encode_rgba8(const uint8_t* rgba, int width, int height, std::vector<uint8_t>& out) {
    out.clear();
    out.insert(out.end(), kSignature, kSignature + 8);
    uint8_t ihdr[13];
    for (int k = 0; k < 4; k++) { ihdr[k] = uint8_t(uint32_t(width) >> (24 - 8 * k)); ihdr[4 + k] = uint8_t(uint32_t(height) >> (24 - 8 * k)); }
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    put_chunk(out, "IHDR", ihdr, 13);
    const size_t rb = size_t(width) * 4, bpp = 4;
    std::vector<uint8_t> filtered(size_t(height) * (rb + 1)), zero(rb, 0), trial[5];
    for (auto& t : trial) t.resize(rb);
    for (int y = 0; y < height; y++) {
        const uint8_t* cur = rgba + size_t(y) * rb;
        const uint8_t* up = y ? cur - rb : zero.data();
        size_t best = 0;
        uint64_t best_cost = ~uint64_t(0);
        for (int ft = 0; ft < 5; ft++) {       // the usual heuristic: the filter whose output, read as signed bytes, is smallest
            uint8_t* t = trial[ft].data();
            uint64_t cost = 0;
            for (size_t i = 0; i < rb; i++) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
                const int pred = ft == 0 ? 0 : ft == 1 ? a : ft == 2 ? b : ft == 3 ? ((a + b) >> 1) : paeth(a, b, c);
                const uint8_t v = uint8_t(cur[i] - pred);
                t[i] = v;
                cost += v < 128 ? v : 256 - v;
            }
            if (cost < best_cost) { best_cost = cost; best = size_t(ft); }
        }
        uint8_t* dst = filtered.data() + size_t(y) * (rb + 1);
        dst[0] = uint8_t(best);
        std::memcpy(dst + 1, trial[best].data(), rb);
    }
    std::vector<uint8_t> z;
    zlib_deflate(filtered.data(), filtered.size(), z);
    for (size_t at = 0; at < z.size(); at += size_t(1) << 18)      // 256 KiB chunks (z is never empty: header + end-of-block + checksum)
        put_chunk(out, "IDAT", z.data() + at, std::min(z.size() - at, size_t(1) << 18));
    put_chunk(out, "IEND", nullptr, 0);
}

}  // namespace ph
