// Host-side audio ingest / egress (include/asr_audio_io.h): native FLAC decoder and
// encoder, 16-bit WAV, and a thread pool that decodes a file list straight into the
// packed int16 batch buffer fe_run() consumes.
//
// Replaces sf.read (/root/reference/preprocess.py:69) and the 16-bit files SoX writes
// for the augmented copies (/root/reference/utils/augmentation.py:28,53).  Written from
// the FLAC format specification (RFC 9639); no code from libFLAC / libsndfile.
#include "../../include/asr_audio_io.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------- CRC / MD5
struct CrcTables {
    uint8_t c8[256];
    uint16_t c16[256];
    CrcTables() {
        for (int i = 0; i < 256; ++i) {
            uint8_t a = (uint8_t)i;
            for (int k = 0; k < 8; ++k) a = (uint8_t)((a & 0x80) ? ((a << 1) ^ 0x07) : (a << 1));
            c8[i] = a;
            uint16_t b = (uint16_t)(i << 8);
            for (int k = 0; k < 8; ++k) b = (uint16_t)((b & 0x8000) ? ((b << 1) ^ 0x8005) : (b << 1));
            c16[i] = b;
        }
    }
};
const CrcTables kCrc;

inline uint8_t crc8(const uint8_t* p, size_t n) {
    uint8_t c = 0;
    for (size_t i = 0; i < n; ++i) c = kCrc.c8[c ^ p[i]];
    return c;
}
inline uint16_t crc16(const uint8_t* p, size_t n) {
    uint16_t c = 0;
    for (size_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ kCrc.c16[(c >> 8) ^ p[i]]);
    return c;
}

struct Md5 {
    uint32_t s[4];
    uint64_t len;
    uint8_t buf[64];
    size_t fill;
    Md5() : len(0), fill(0) {
        s[0] = 0x67452301u; s[1] = 0xefcdab89u; s[2] = 0x98badcfeu; s[3] = 0x10325476u;
    }
    static inline uint32_t rol(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }
    void block(const uint8_t* p) {
        uint32_t w[16];
        for (int i = 0; i < 16; ++i)
            w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                   ((uint32_t)p[4 * i + 3] << 24);
        uint32_t a = s[0], b = s[1], c = s[2], d = s[3];
#define MD5_F(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define MD5_G(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define MD5_STEP(f, a, b, c, d, g, k, r) a += f(b, c, d) + w[g] + k; a = rol(a, r) + b;
        MD5_STEP(MD5_F, a, b, c, d, 0, 0xd76aa478u, 7)   MD5_STEP(MD5_F, d, a, b, c, 1, 0xe8c7b756u, 12)
        MD5_STEP(MD5_F, c, d, a, b, 2, 0x242070dbu, 17)  MD5_STEP(MD5_F, b, c, d, a, 3, 0xc1bdceeeu, 22)
        MD5_STEP(MD5_F, a, b, c, d, 4, 0xf57c0fafu, 7)   MD5_STEP(MD5_F, d, a, b, c, 5, 0x4787c62au, 12)
        MD5_STEP(MD5_F, c, d, a, b, 6, 0xa8304613u, 17)  MD5_STEP(MD5_F, b, c, d, a, 7, 0xfd469501u, 22)
        MD5_STEP(MD5_F, a, b, c, d, 8, 0x698098d8u, 7)   MD5_STEP(MD5_F, d, a, b, c, 9, 0x8b44f7afu, 12)
        MD5_STEP(MD5_F, c, d, a, b, 10, 0xffff5bb1u, 17) MD5_STEP(MD5_F, b, c, d, a, 11, 0x895cd7beu, 22)
        MD5_STEP(MD5_F, a, b, c, d, 12, 0x6b901122u, 7)  MD5_STEP(MD5_F, d, a, b, c, 13, 0xfd987193u, 12)
        MD5_STEP(MD5_F, c, d, a, b, 14, 0xa679438eu, 17) MD5_STEP(MD5_F, b, c, d, a, 15, 0x49b40821u, 22)
        MD5_STEP(MD5_G, a, b, c, d, 1, 0xf61e2562u, 5)   MD5_STEP(MD5_G, d, a, b, c, 6, 0xc040b340u, 9)
        MD5_STEP(MD5_G, c, d, a, b, 11, 0x265e5a51u, 14) MD5_STEP(MD5_G, b, c, d, a, 0, 0xe9b6c7aau, 20)
        MD5_STEP(MD5_G, a, b, c, d, 5, 0xd62f105du, 5)   MD5_STEP(MD5_G, d, a, b, c, 10, 0x02441453u, 9)
        MD5_STEP(MD5_G, c, d, a, b, 15, 0xd8a1e681u, 14) MD5_STEP(MD5_G, b, c, d, a, 4, 0xe7d3fbc8u, 20)
        MD5_STEP(MD5_G, a, b, c, d, 9, 0x21e1cde6u, 5)   MD5_STEP(MD5_G, d, a, b, c, 14, 0xc33707d6u, 9)
        MD5_STEP(MD5_G, c, d, a, b, 3, 0xf4d50d87u, 14)  MD5_STEP(MD5_G, b, c, d, a, 8, 0x455a14edu, 20)
        MD5_STEP(MD5_G, a, b, c, d, 13, 0xa9e3e905u, 5)  MD5_STEP(MD5_G, d, a, b, c, 2, 0xfcefa3f8u, 9)
        MD5_STEP(MD5_G, c, d, a, b, 7, 0x676f02d9u, 14)  MD5_STEP(MD5_G, b, c, d, a, 12, 0x8d2a4c8au, 20)
        MD5_STEP(MD5_H, a, b, c, d, 5, 0xfffa3942u, 4)   MD5_STEP(MD5_H, d, a, b, c, 8, 0x8771f681u, 11)
        MD5_STEP(MD5_H, c, d, a, b, 11, 0x6d9d6122u, 16) MD5_STEP(MD5_H, b, c, d, a, 14, 0xfde5380cu, 23)
        MD5_STEP(MD5_H, a, b, c, d, 1, 0xa4beea44u, 4)   MD5_STEP(MD5_H, d, a, b, c, 4, 0x4bdecfa9u, 11)
        MD5_STEP(MD5_H, c, d, a, b, 7, 0xf6bb4b60u, 16)  MD5_STEP(MD5_H, b, c, d, a, 10, 0xbebfbc70u, 23)
        MD5_STEP(MD5_H, a, b, c, d, 13, 0x289b7ec6u, 4)  MD5_STEP(MD5_H, d, a, b, c, 0, 0xeaa127fau, 11)
        MD5_STEP(MD5_H, c, d, a, b, 3, 0xd4ef3085u, 16)  MD5_STEP(MD5_H, b, c, d, a, 6, 0x04881d05u, 23)
        MD5_STEP(MD5_H, a, b, c, d, 9, 0xd9d4d039u, 4)   MD5_STEP(MD5_H, d, a, b, c, 12, 0xe6db99e5u, 11)
        MD5_STEP(MD5_H, c, d, a, b, 15, 0x1fa27cf8u, 16) MD5_STEP(MD5_H, b, c, d, a, 2, 0xc4ac5665u, 23)
        MD5_STEP(MD5_I, a, b, c, d, 0, 0xf4292244u, 6)   MD5_STEP(MD5_I, d, a, b, c, 7, 0x432aff97u, 10)
        MD5_STEP(MD5_I, c, d, a, b, 14, 0xab9423a7u, 15) MD5_STEP(MD5_I, b, c, d, a, 5, 0xfc93a039u, 21)
        MD5_STEP(MD5_I, a, b, c, d, 12, 0x655b59c3u, 6)  MD5_STEP(MD5_I, d, a, b, c, 3, 0x8f0ccc92u, 10)
        MD5_STEP(MD5_I, c, d, a, b, 10, 0xffeff47du, 15) MD5_STEP(MD5_I, b, c, d, a, 1, 0x85845dd1u, 21)
        MD5_STEP(MD5_I, a, b, c, d, 8, 0x6fa87e4fu, 6)   MD5_STEP(MD5_I, d, a, b, c, 15, 0xfe2ce6e0u, 10)
        MD5_STEP(MD5_I, c, d, a, b, 6, 0xa3014314u, 15)  MD5_STEP(MD5_I, b, c, d, a, 13, 0x4e0811a1u, 21)
        MD5_STEP(MD5_I, a, b, c, d, 4, 0xf7537e82u, 6)   MD5_STEP(MD5_I, d, a, b, c, 11, 0xbd3af235u, 10)
        MD5_STEP(MD5_I, c, d, a, b, 2, 0x2ad7d2bbu, 15)  MD5_STEP(MD5_I, b, c, d, a, 9, 0xeb86d391u, 21)
#undef MD5_STEP
#undef MD5_F
#undef MD5_G
#undef MD5_H
#undef MD5_I
        s[0] += a; s[1] += b; s[2] += c; s[3] += d;
    }
    void update(const uint8_t* p, size_t n) {
        len += n;
        if (fill) {
            size_t take = 64 - fill < n ? 64 - fill : n;
            memcpy(buf + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) { block(buf); fill = 0; }
        }
        while (n >= 64) { block(p); p += 64; n -= 64; }
        if (n) { memcpy(buf, p, n); fill = n; }
    }
    void finish(uint8_t out[16]) {
        uint64_t bits = len * 8;
        uint8_t pad[72];
        size_t padn = (fill < 56) ? 56 - fill : 120 - fill;
        memset(pad, 0, sizeof(pad));
        pad[0] = 0x80;
        for (int i = 0; i < 8; ++i) pad[padn + i] = (uint8_t)(bits >> (8 * i));
        update(pad, padn + 8);
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 4; ++k) out[4 * i + k] = (uint8_t)(s[i] >> (8 * k));
    }
};

const uint16_t kEndianProbe = 1;
const bool kLittleEndian = *(const uint8_t*)&kEndianProbe == 1;

// ----------------------------------------------------------------------------- bit reader
struct BitReader {
    const uint8_t* base;
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc;   // valid bits are the top `cnt`; everything below is zero
    int cnt;
    bool bad;
    void init(const uint8_t* b, const uint8_t* e) { base = p = b; end = e; acc = 0; cnt = 0; bad = false; }
    inline void refill() {
        while (cnt <= 56 && p < end) {
            acc |= (uint64_t)(*p++) << (56 - cnt);
            cnt += 8;
        }
    }
    inline uint32_t get(int n) {   // 0 <= n <= 32
        if (n == 0) return 0;
        if (cnt < n) {
            refill();
            if (cnt < n) { bad = true; cnt = 0; acc = 0; return 0; }
        }
        uint32_t v = (uint32_t)(acc >> (64 - n));
        acc <<= n;
        cnt -= n;
        return v;
    }
    inline int32_t get_signed(int n) {   // 1 <= n <= 32
        uint32_t v = get(n);
        if (n < 32) {
            uint32_t m = 1u << (n - 1);
            return (int32_t)((v ^ m) - m);
        }
        return (int32_t)v;
    }
    inline uint32_t unary() {   // number of 0 bits before the next 1 bit (the 1 is consumed)
        uint32_t q = 0;
        for (;;) {
            if (cnt == 0 || acc == 0) {
                q += (uint32_t)cnt;
                acc = 0;
                cnt = 0;
                refill();
                if (cnt == 0) { bad = true; return 0; }
                if (acc == 0) continue;
            }
            int z = __builtin_clzll(acc);
            q += (uint32_t)z;
            acc = (z == 63) ? 0 : acc << (z + 1);
            cnt -= z + 1;
            return q;
        }
    }
    inline void align() {
        int r = cnt & 7;
        if (r) { acc <<= r; cnt -= r; }
    }
    inline size_t byte_pos() const { return (size_t)(p - base) - (size_t)(cnt >> 3); }   // call when aligned
};

// ----------------------------------------------------------------------------- FLAC stream header
struct StreamInfo {
    int min_block, max_block;
    int sample_rate, channels, bps;
    int64_t total;
    uint8_t md5[16];
    bool has_md5;
    size_t first_frame;   // byte offset of the first audio frame
};

int parse_flac_header(const uint8_t* d, int64_t n, StreamInfo* si) {
    size_t pos = 0;
    if (n >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3') {   // ID3v2 tag in front of the stream
        size_t sz = ((size_t)(d[6] & 0x7f) << 21) | ((size_t)(d[7] & 0x7f) << 14) | ((size_t)(d[8] & 0x7f) << 7) |
                    (size_t)(d[9] & 0x7f);
        pos = 10 + sz;
    }
    if ((int64_t)pos + 4 > n || memcmp(d + pos, "fLaC", 4) != 0) return AIO_ERR_FORMAT;
    pos += 4;
    bool seen = false;
    for (;;) {
        if ((int64_t)pos + 4 > n) return AIO_ERR_FORMAT;
        int last = d[pos] >> 7, type = d[pos] & 0x7f;
        size_t len = ((size_t)d[pos + 1] << 16) | ((size_t)d[pos + 2] << 8) | d[pos + 3];
        pos += 4;
        if ((int64_t)(pos + len) > n) return AIO_ERR_FORMAT;
        if (type == 0) {
            if (len < 34) return AIO_ERR_FORMAT;
            const uint8_t* s = d + pos;
            si->min_block = (s[0] << 8) | s[1];
            si->max_block = (s[2] << 8) | s[3];
            si->sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
            si->channels = ((s[12] >> 1) & 7) + 1;
            si->bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            si->total = ((int64_t)(s[13] & 0x0f) << 32) | ((int64_t)s[14] << 24) | ((int64_t)s[15] << 16) |
                        ((int64_t)s[16] << 8) | (int64_t)s[17];
            memcpy(si->md5, s + 18, 16);
            si->has_md5 = false;
            for (int i = 0; i < 16; ++i) si->has_md5 |= (s[18 + i] != 0);
            seen = true;
        } else if (type == 127) {
            return AIO_ERR_FORMAT;
        }
        pos += len;
        if (last) break;
    }
    if (!seen || si->sample_rate == 0) return AIO_ERR_FORMAT;
    si->first_frame = pos;
    return AIO_OK;
}

// ----------------------------------------------------------------------------- FLAC frame decode
struct FrameHeader {
    int blocksize, sample_rate, chan_assign, channels, bps;
};

int read_frame_header(BitReader& br, const StreamInfo& si, FrameHeader* fh) {
    size_t start = br.byte_pos();
    uint32_t sync = br.get(15);
    if (br.bad || sync != 0x7ffc) return AIO_ERR_FORMAT;   // 11111111 111110 + reserved 0
    br.get(1);                                             // blocking strategy (not needed to decode)
    int bs_code = (int)br.get(4), sr_code = (int)br.get(4);
    int ch_code = (int)br.get(4), ss_code = (int)br.get(3);
    if (br.get(1) != 0) return AIO_ERR_FORMAT;
    // UTF-8 style coded frame / sample number (1..7 bytes)
    uint32_t b0 = br.get(8);
    int extra = 0;
    if (b0 & 0x80) {
        if ((b0 & 0xe0) == 0xc0) extra = 1;
        else if ((b0 & 0xf0) == 0xe0) extra = 2;
        else if ((b0 & 0xf8) == 0xf0) extra = 3;
        else if ((b0 & 0xfc) == 0xf8) extra = 4;
        else if ((b0 & 0xfe) == 0xfc) extra = 5;
        else if (b0 == 0xfe) extra = 6;
        else return AIO_ERR_FORMAT;
    }
    for (int i = 0; i < extra; ++i)
        if ((br.get(8) & 0xc0) != 0x80) return AIO_ERR_FORMAT;
    int bs;
    if (bs_code == 0) return AIO_ERR_FORMAT;
    else if (bs_code == 1) bs = 192;
    else if (bs_code <= 5) bs = 576 << (bs_code - 2);
    else if (bs_code == 6) bs = (int)br.get(8) + 1;
    else if (bs_code == 7) bs = (int)br.get(16) + 1;
    else bs = 256 << (bs_code - 8);
    static const int kRates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    int sr;
    if (sr_code < 12) sr = sr_code ? kRates[sr_code] : si.sample_rate;
    else if (sr_code == 12) sr = (int)br.get(8) * 1000;
    else if (sr_code == 13) sr = (int)br.get(16);
    else if (sr_code == 14) sr = (int)br.get(16) * 10;
    else return AIO_ERR_FORMAT;
    static const int kBits[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    int bps = kBits[ss_code];
    if (bps < 0) return AIO_ERR_FORMAT;
    if (bps == 0) bps = si.bps;
    if (ch_code > 10) return AIO_ERR_FORMAT;
    fh->blocksize = bs;
    fh->sample_rate = sr;
    fh->chan_assign = ch_code;
    fh->channels = ch_code < 8 ? ch_code + 1 : 2;
    fh->bps = bps;
    if (br.bad) return AIO_ERR_FORMAT;
    size_t hdr_end = br.byte_pos();
    uint32_t want = br.get(8);
    if (br.bad || crc8(br.base + start, hdr_end - start) != want) return AIO_ERR_FORMAT;
    return AIO_OK;
}

// One Rice partition of n residuals with parameter k, on a local copy of the bit window.  The window is topped up
// eight bytes at a time (bits below `cnt` may then hold look-ahead data: they are the true next bits, OR-ing them
// again on the next top-up is idempotent); a code that does not fit the window, and the last bytes of the
// stream, go through the generic reader.
inline void rice_partition(BitReader& br, int k, int n, int32_t* res) {
    uint64_t acc = br.acc;
    int cnt = br.cnt;
    const uint8_t* p = br.p;
    const uint8_t* const end = br.end;
    int i = 0;
    for (; i < n; ++i) {
        if (cnt < 48) {
            if (end - p < 8) break;
            uint64_t w;
            memcpy(&w, p, 8);
            w = __builtin_bswap64(w);
            acc |= w >> cnt;
            p += (63 - cnt) >> 3;
            cnt |= 56;
        }
        const uint64_t valid = acc & (~0ull << (64 - cnt));           // cnt >= 48 here
        if (valid == 0) break;
        const int z = __builtin_clzll(valid);
        if (z + 1 + k > cnt) break;
        acc <<= z + 1;                                                // z + 1 <= 48
        cnt -= z + 1;
        uint32_t u = (uint32_t)z << k;
        if (k) {
            u |= (uint32_t)(acc >> (64 - k));
            acc <<= k;
            cnt -= k;
        }
        res[i] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
    }
    br.acc = cnt ? acc & (~0ull << (64 - cnt)) : 0;                   // restore the reader's invariant: zeros below cnt
    br.cnt = cnt;
    br.p = p;
    for (; i < n; ++i) {                                              // generic path: window edge cases, end of stream
        uint32_t hi = br.unary();
        uint32_t u = (hi << k) | br.get(k);
        res[i] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
        if (br.bad) return;
    }
}

// Prediction recurrence out[i] = r[i - ORDER] + (sum_j coef[j] out[i-1-j] >> shift) with a compile-time order.
// ACC = int32_t when the sums fit 32 bits (bps + precision + ceil(log2 order) <= 32), else int64_t.
template <int ORDER, class ACC>
void lpc_restore(const int32_t* r, int n, const int32_t* coef, int shift, int32_t* out) {
    ACC c[ORDER];
    for (int j = 0; j < ORDER; ++j) c[j] = coef[j];
    for (int i = ORDER; i < n; ++i) {
        ACC sum = 0;
        for (int j = 0; j < ORDER; ++j) sum += c[j] * (ACC)out[i - 1 - j];
        out[i] = (int32_t)((ACC)r[i - ORDER] + (sum >> shift));
    }
}

template <class ACC>
bool lpc_restore_dispatch(int order, const int32_t* r, int n, const int32_t* coef, int shift, int32_t* out) {
    switch (order) {
        case 1: lpc_restore<1, ACC>(r, n, coef, shift, out); return true;
        case 2: lpc_restore<2, ACC>(r, n, coef, shift, out); return true;
        case 3: lpc_restore<3, ACC>(r, n, coef, shift, out); return true;
        case 4: lpc_restore<4, ACC>(r, n, coef, shift, out); return true;
        case 5: lpc_restore<5, ACC>(r, n, coef, shift, out); return true;
        case 6: lpc_restore<6, ACC>(r, n, coef, shift, out); return true;
        case 7: lpc_restore<7, ACC>(r, n, coef, shift, out); return true;
        case 8: lpc_restore<8, ACC>(r, n, coef, shift, out); return true;
        case 9: lpc_restore<9, ACC>(r, n, coef, shift, out); return true;
        case 10: lpc_restore<10, ACC>(r, n, coef, shift, out); return true;
        case 11: lpc_restore<11, ACC>(r, n, coef, shift, out); return true;
        case 12: lpc_restore<12, ACC>(r, n, coef, shift, out); return true;
        default: return false;
    }
}

int read_residual(BitReader& br, int blocksize, int order, int32_t* res) {
    int method = (int)br.get(2);
    if (method > 1) return AIO_ERR_FORMAT;
    int pbits = method ? 5 : 4, esc = method ? 31 : 15;
    int porder = (int)br.get(4);
    int parts = 1 << porder;
    if ((blocksize >> porder) << porder != blocksize && porder > 0) return AIO_ERR_FORMAT;
    int per = blocksize >> porder;
    if (per < order && porder > 0) return AIO_ERR_FORMAT;
    int idx = 0;   // index into res[] (res holds blocksize - order values)
    for (int q = 0; q < parts; ++q) {
        int n = (q == 0) ? per - order : per;
        if (n < 0) return AIO_ERR_FORMAT;
        int k = (int)br.get(pbits);
        if (k == esc) {
            int raw = (int)br.get(5);
            for (int i = 0; i < n; ++i) res[idx++] = raw ? br.get_signed(raw) : 0;
        } else {
            rice_partition(br, k, n, res + idx);
            idx += n;
        }
        if (br.bad) return AIO_ERR_FORMAT;
    }
    return AIO_OK;
}

int read_subframe(BitReader& br, int blocksize, int bps, int32_t* out, std::vector<int32_t>& scratch) {
    if (br.get(1) != 0) return AIO_ERR_FORMAT;
    int type = (int)br.get(6);
    int wasted = 0;
    if (br.get(1)) wasted = (int)br.unary() + 1;
    if (br.bad || wasted >= bps) return AIO_ERR_FORMAT;
    bps -= wasted;
    if (type == 0) {                                        // CONSTANT
        int32_t v = br.get_signed(bps);
        for (int i = 0; i < blocksize; ++i) out[i] = v;
    } else if (type == 1) {                                 // VERBATIM
        for (int i = 0; i < blocksize; ++i) out[i] = br.get_signed(bps);
    } else if (type >= 8 && type <= 12) {                   // FIXED, order 0..4
        int order = type - 8;
        if (order > blocksize) return AIO_ERR_FORMAT;
        for (int i = 0; i < order; ++i) out[i] = br.get_signed(bps);
        scratch.resize((size_t)blocksize);
        int rc = read_residual(br, blocksize, order, scratch.data());
        if (rc) return rc;
        const int32_t* r = scratch.data();
        // 64-bit accumulators: a 17-bit side channel at order 4 overflows 32 bits only in
        // invalid streams, but wrap-around must not be undefined behaviour either way.
        switch (order) {
            case 0: for (int i = 0; i < blocksize; ++i) out[i] = r[i]; break;
            case 1: for (int i = 1; i < blocksize; ++i) out[i] = (int32_t)((int64_t)r[i - 1] + out[i - 1]); break;
            case 2: for (int i = 2; i < blocksize; ++i)
                        out[i] = (int32_t)((int64_t)r[i - 2] + 2 * (int64_t)out[i - 1] - out[i - 2]);
                    break;
            case 3: for (int i = 3; i < blocksize; ++i)
                        out[i] = (int32_t)((int64_t)r[i - 3] + 3 * (int64_t)out[i - 1] - 3 * (int64_t)out[i - 2] + out[i - 3]);
                    break;
            default: for (int i = 4; i < blocksize; ++i)
                        out[i] = (int32_t)((int64_t)r[i - 4] + 4 * (int64_t)out[i - 1] - 6 * (int64_t)out[i - 2] +
                                           4 * (int64_t)out[i - 3] - out[i - 4]);
        }
    } else if (type >= 32) {                                // LPC, order 1..32
        int order = type - 31;
        if (order > blocksize) return AIO_ERR_FORMAT;
        for (int i = 0; i < order; ++i) out[i] = br.get_signed(bps);
        int prec = (int)br.get(4) + 1;
        if (prec == 16) return AIO_ERR_FORMAT;
        int shift = br.get_signed(5);
        if (shift < 0) return AIO_ERR_FORMAT;
        int32_t coef[32];
        for (int i = 0; i < order; ++i) coef[i] = br.get_signed(prec);
        scratch.resize((size_t)blocksize);
        int rc = read_residual(br, blocksize, order, scratch.data());
        if (rc) return rc;
        const int32_t* r = scratch.data();
        int lg = 0;
        while ((1 << lg) < order) ++lg;
        const bool narrow = bps + prec + lg <= 32;                   // every libFLAC stream at <= 16 bits
        const bool done = narrow ? lpc_restore_dispatch<int32_t>(order, r, blocksize, coef, shift, out)
                                 : lpc_restore_dispatch<int64_t>(order, r, blocksize, coef, shift, out);
        if (!done) {
            for (int i = order; i < blocksize; ++i) {
                int64_t acc = 0;
                for (int j = 0; j < order; ++j) acc += (int64_t)coef[j] * out[i - 1 - j];
                out[i] = (int32_t)((int64_t)r[i - order] + (acc >> shift));
            }
        }
    } else {
        return AIO_ERR_FORMAT;                              // reserved subframe type
    }
    if (br.bad) return AIO_ERR_FORMAT;
    if (wasted)
        for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)((uint32_t)out[i] << wasted);
    return AIO_OK;
}

int decode_flac(const uint8_t* d, int64_t n, int16_t* out, int64_t capacity, int64_t* n_decoded, int check_md5) {
    StreamInfo si;
    int rc = parse_flac_header(d, n, &si);
    if (rc) return rc;
    if (si.bps > 16 || si.bps < 4) return AIO_ERR_UNSUPPORTED;
    const int C = si.channels, up = 16 - si.bps;
    BitReader br;
    br.init(d + si.first_frame, d + n);
    std::vector<int32_t> chan[8], scratch;
    Md5 md5;
    std::vector<uint8_t> md5buf;
    const bool do_md5 = check_md5 && si.has_md5;
    int64_t done = 0;
    for (;;) {
        br.refill();
        if (br.cnt == 0) break;                                        // clean end of stream
        if (si.total > 0 && done >= si.total) break;                   // trailing bytes (e.g. ID3v1) after the audio
        size_t frame_start = br.byte_pos();
        FrameHeader fh;
        rc = read_frame_header(br, si, &fh);
        if (rc) return rc;
        if (fh.channels != C || fh.bps != si.bps || fh.blocksize <= 0) return AIO_ERR_FORMAT;
        for (int c = 0; c < C; ++c) {
            chan[c].resize((size_t)fh.blocksize);
            int bps = fh.bps;
            if ((fh.chan_assign == 8 && c == 1) || (fh.chan_assign == 9 && c == 0) || (fh.chan_assign == 10 && c == 1))
                bps += 1;                                              // the side channel carries one more bit
            rc = read_subframe(br, fh.blocksize, bps, chan[c].data(), scratch);
            if (rc) return rc;
        }
        br.align();
        size_t body_end = br.byte_pos();
        uint32_t want = br.get(16);
        if (br.bad || crc16(br.base + frame_start, body_end - frame_start) != want) return AIO_ERR_FORMAT;
        const int bs = fh.blocksize;
        if (fh.chan_assign == 8) {
            for (int i = 0; i < bs; ++i) chan[1][i] = chan[0][i] - chan[1][i];
        } else if (fh.chan_assign == 9) {
            for (int i = 0; i < bs; ++i) chan[0][i] = chan[0][i] + chan[1][i];
        } else if (fh.chan_assign == 10) {
            for (int i = 0; i < bs; ++i) {
                int32_t side = chan[1][i];
                int32_t mid = (int32_t)(((uint32_t)chan[0][i] << 1) | (uint32_t)(side & 1));
                chan[0][i] = (mid + side) >> 1;
                chan[1][i] = (mid - side) >> 1;
            }
        }
        if (do_md5 && !(C == 1 && up == 0 && kLittleEndian)) {         // MD5 is over the un-scaled samples
            const int bytes = (si.bps + 7) / 8;
            md5buf.resize((size_t)bs * C * bytes);
            uint8_t* m = md5buf.data();
            for (int i = 0; i < bs; ++i)
                for (int c = 0; c < C; ++c) {
                    uint32_t v = (uint32_t)chan[c][i];
                    for (int b = 0; b < bytes; ++b) *m++ = (uint8_t)(v >> (8 * b));
                }
            md5.update(md5buf.data(), md5buf.size());
        }
        if ((done + bs) * C > capacity) return AIO_ERR_CAPACITY;
        int16_t* o = out + done * C;
        if (C == 1) {
            const int32_t* s = chan[0].data();
            for (int i = 0; i < bs; ++i) o[i] = (int16_t)(s[i] * (1 << up));
            if (do_md5 && up == 0 && kLittleEndian) md5.update((const uint8_t*)o, (size_t)bs * 2);   // 16-bit mono: the output IS the hashed byte stream
        } else {
            for (int i = 0; i < bs; ++i)
                for (int c = 0; c < C; ++c) o[(size_t)i * C + c] = (int16_t)(chan[c][i] * (1 << up));
        }
        done += bs;
    }
    if (si.total > 0 && done != si.total) return AIO_ERR_FORMAT;       // truncated stream
    if (do_md5) {
        uint8_t got[16];
        md5.finish(got);
        if (memcmp(got, si.md5, 16) != 0) return AIO_ERR_FORMAT;
    }
    *n_decoded = done;
    return AIO_OK;
}

// FLAC streams without total_samples (streamed encodes): count by walking the frame headers is
// not possible without decoding the subframes, so probe reports -1 and the caller decodes with
// a generous capacity.

// ----------------------------------------------------------------------------- WAV
struct WavInfo {
    int sample_rate, channels, bps, format_tag;
    size_t data_off, data_len;
};

inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint32_t le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

int parse_wav(const uint8_t* d, int64_t n, WavInfo* w) {
    if (n < 12 || memcmp(d, "RIFF", 4) != 0 || memcmp(d + 8, "WAVE", 4) != 0) return AIO_ERR_FORMAT;
    size_t pos = 12;
    bool fmt = false;
    while ((int64_t)pos + 8 <= n) {
        uint32_t len = le32(d + pos + 4);
        const uint8_t* body = d + pos + 8;
        if (memcmp(d + pos, "fmt ", 4) == 0) {
            if (len < 16 || (int64_t)(pos + 8 + len) > n) return AIO_ERR_FORMAT;
            w->format_tag = (int)le16(body);
            w->channels = (int)le16(body + 2);
            w->sample_rate = (int)le32(body + 4);
            w->bps = (int)le16(body + 14);
            if (w->format_tag == 0xfffe && len >= 26) w->format_tag = (int)le16(body + 24);   // WAVE_FORMAT_EXTENSIBLE
            fmt = true;
        } else if (memcmp(d + pos, "data", 4) == 0) {
            if (!fmt) return AIO_ERR_FORMAT;
            w->data_off = pos + 8;
            size_t avail = (size_t)n - w->data_off;
            w->data_len = len < avail ? len : avail;   // tolerate a size field longer than the file
            return AIO_OK;
        }
        pos += 8 + (size_t)len + (len & 1);
    }
    return AIO_ERR_FORMAT;
}

int decode_wav(const uint8_t* d, int64_t n, int16_t* out, int64_t capacity, int64_t* n_decoded) {
    WavInfo w;
    int rc = parse_wav(d, n, &w);
    if (rc) return rc;
    if (w.format_tag != 1 || w.bps != 16 || w.channels < 1) return AIO_ERR_UNSUPPORTED;
    int64_t total = (int64_t)(w.data_len / 2);
    if (total > capacity) return AIO_ERR_CAPACITY;
    const uint8_t* s = d + w.data_off;
    for (int64_t i = 0; i < total; ++i) out[i] = (int16_t)le16(s + 2 * i);
    *n_decoded = total / w.channels;
    return AIO_OK;
}

// ----------------------------------------------------------------------------- files
int read_whole_file(const char* path, std::vector<uint8_t>& buf, size_t limit = 0) {
    FILE* f = fopen(path, "rb");
    if (!f) return AIO_ERR_IO;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return AIO_ERR_IO; }
    long sz = ftell(f);
    if (sz < 0) { fclose(f); return AIO_ERR_IO; }
    rewind(f);
    size_t want = (size_t)sz;
    if (limit && want > limit) want = limit;
    buf.resize(want);
    size_t got = want ? fread(buf.data(), 1, want, f) : 0;
    fclose(f);
    return got == want ? AIO_OK : AIO_ERR_IO;
}

template <class F>
void parallel_for(int32_t n, int32_t n_threads, F fn) {
    if (n_threads <= 0) n_threads = (int32_t)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = n;
    if (n_threads <= 1) {
        for (int32_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int32_t> next(0);
    std::vector<std::thread> pool;
    pool.reserve((size_t)n_threads);
    for (int32_t t = 0; t < n_threads; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                int32_t i = next.fetch_add(1);
                if (i >= n) return;
                fn(i);
            }
        });
    for (auto& th : pool) th.join();
}

// ----------------------------------------------------------------------------- FLAC encoder
struct BitWriter {
    std::vector<uint8_t>& b;
    uint64_t acc;
    int cnt;
    explicit BitWriter(std::vector<uint8_t>& buf) : b(buf), acc(0), cnt(0) {}
    inline void put(uint32_t v, int n) {   // 0 <= n <= 32
        if (n == 0) return;
        if (n < 32) v &= (1u << n) - 1;
        acc = (acc << n) | v;
        cnt += n;
        while (cnt >= 8) {
            b.push_back((uint8_t)(acc >> (cnt - 8)));
            cnt -= 8;
        }
    }
    inline void unary(uint32_t q) {        // q zeros then a one
        while (q >= 32) { put(0, 32); q -= 32; }
        put(1, (int)q + 1);
    }
    inline void align() { if (cnt) put(0, 8 - cnt); }
};

const int kBlock = 4096;

// bits needed for the Rice-coded residual with the best partition order; fills the plan
struct RicePlan {
    int porder;
    std::vector<int> k;
    uint64_t bits;
};

void plan_rice(const uint32_t* u, int n_res, int order, int blocksize, RicePlan* best) {
    best->bits = ~0ull;
    int max_po = 0;
    while (max_po < 8 && (blocksize & ((1 << (max_po + 1)) - 1)) == 0 && (blocksize >> (max_po + 1)) > order) ++max_po;
    std::vector<int> ks;
    for (int po = 0; po <= max_po; ++po) {
        int parts = 1 << po, per = blocksize >> po;
        uint64_t bits = 6;
        ks.assign((size_t)parts, 0);
        int idx = 0;
        for (int q = 0; q < parts; ++q) {
            int n = q == 0 ? per - order : per;
            uint64_t sum = 0;
            for (int i = 0; i < n; ++i) sum += u[idx + i];
            // start near log2(mean) and look at the neighbours
            int k0 = 0;
            if (n > 0) { uint64_t mean = sum / (uint64_t)n; while (k0 < 14 && (mean >> (k0 + 1))) ++k0; }
            uint64_t bb = ~0ull;
            int bk = 0;
            for (int k = (k0 > 0 ? k0 - 1 : 0); k <= (k0 < 14 ? k0 + 1 : 14); ++k) {
                uint64_t t = (uint64_t)n * (uint64_t)(k + 1);
                for (int i = 0; i < n; ++i) t += u[idx + i] >> k;
                if (t < bb) { bb = t; bk = k; }
            }
            ks[(size_t)q] = bk;
            bits += 4 + bb;
            idx += n;
        }
        (void)n_res;
        if (bits < best->bits) { best->bits = bits; best->porder = po; best->k = ks; }
    }
}

void put_utf8(BitWriter& bw, uint64_t v) {
    if (v < 0x80) { bw.put((uint32_t)v, 8); return; }
    int extra = v < 0x800 ? 1 : v < 0x10000 ? 2 : v < 0x200000 ? 3 : v < 0x4000000 ? 4 : v < 0x80000000ull ? 5 : 6;
    uint32_t lead = (0xff00u >> (extra + 1)) & 0xff;   // extra+1 leading ones
    int lead_bits = 6 - extra;                         // payload bits in the first byte
    bw.put(lead | (lead_bits > 0 ? (uint32_t)(v >> (6 * extra)) & ((1u << lead_bits) - 1) : 0), 8);
    for (int i = extra - 1; i >= 0; --i) bw.put(0x80u | (uint32_t)((v >> (6 * i)) & 0x3f), 8);
}

void encode_subframe(BitWriter& bw, const int32_t* s, int n, int bps, std::vector<int32_t>& res,
                     std::vector<uint32_t>& zz) {
    bool constant = true;
    for (int i = 1; i < n && constant; ++i) constant = (s[i] == s[0]);
    if (constant) {
        bw.put(0, 1); bw.put(0, 6); bw.put(0, 1);
        bw.put((uint32_t)s[0], bps);
        return;
    }
    // fixed predictors: pick the order with the smallest sum of |residual|
    int best_order = -1;
    uint64_t best_abs = ~0ull;
    int max_order = n > 4 ? 4 : n - 1;
    for (int order = 0; order <= max_order; ++order) {
        uint64_t a = 0;
        for (int i = order; i < n; ++i) {
            int64_t r;
            switch (order) {
                case 0: r = s[i]; break;
                case 1: r = (int64_t)s[i] - s[i - 1]; break;
                case 2: r = (int64_t)s[i] - 2 * (int64_t)s[i - 1] + s[i - 2]; break;
                case 3: r = (int64_t)s[i] - 3 * (int64_t)s[i - 1] + 3 * (int64_t)s[i - 2] - s[i - 3]; break;
                default: r = (int64_t)s[i] - 4 * (int64_t)s[i - 1] + 6 * (int64_t)s[i - 2] - 4 * (int64_t)s[i - 3] + s[i - 4];
            }
            a += (uint64_t)(r < 0 ? -r : r);
        }
        if (a < best_abs) { best_abs = a; best_order = order; }
    }
    const int order = best_order;
    res.resize((size_t)n);
    zz.resize((size_t)n);
    int nres = n - order;
    for (int i = order; i < n; ++i) {
        int64_t r;
        switch (order) {
            case 0: r = s[i]; break;
            case 1: r = (int64_t)s[i] - s[i - 1]; break;
            case 2: r = (int64_t)s[i] - 2 * (int64_t)s[i - 1] + s[i - 2]; break;
            case 3: r = (int64_t)s[i] - 3 * (int64_t)s[i - 1] + 3 * (int64_t)s[i - 2] - s[i - 3]; break;
            default: r = (int64_t)s[i] - 4 * (int64_t)s[i - 1] + 6 * (int64_t)s[i - 2] - 4 * (int64_t)s[i - 3] + s[i - 4];
        }
        res[(size_t)(i - order)] = (int32_t)r;
        zz[(size_t)(i - order)] = ((uint32_t)((int32_t)r) << 1) ^ (uint32_t)((int32_t)r >> 31);
    }
    RicePlan plan;
    plan_rice(zz.data(), nres, order, n, &plan);
    uint64_t fixed_bits = (uint64_t)order * (uint64_t)bps + plan.bits;
    if (fixed_bits >= (uint64_t)n * (uint64_t)bps) {       // incompressible: VERBATIM
        bw.put(0, 1); bw.put(1, 6); bw.put(0, 1);
        for (int i = 0; i < n; ++i) bw.put((uint32_t)s[i], bps);
        return;
    }
    bw.put(0, 1); bw.put((uint32_t)(8 + order), 6); bw.put(0, 1);
    for (int i = 0; i < order; ++i) bw.put((uint32_t)s[i], bps);
    bw.put(0, 2);                                           // Rice, 4-bit parameters
    bw.put((uint32_t)plan.porder, 4);
    int parts = 1 << plan.porder, per = n >> plan.porder, idx = 0;
    for (int q = 0; q < parts; ++q) {
        int cntq = q == 0 ? per - order : per;
        int k = plan.k[(size_t)q];
        bw.put((uint32_t)k, 4);
        for (int i = 0; i < cntq; ++i) {
            uint32_t u = zz[(size_t)(idx + i)];
            bw.unary(u >> k);
            bw.put(u, k);
        }
        idx += cntq;
    }
}

int encode_flac(const int16_t* pcm, int64_t n, int C, int fs, std::vector<uint8_t>& out) {
    if (C < 1 || C > 8 || fs <= 0 || fs >= (1 << 20) || n < 0) return AIO_ERR_INVALID;
    out.clear();
    out.reserve((size_t)(n * C * 2 * 3 / 4 + 256));
    const uint8_t magic[4] = {'f', 'L', 'a', 'C'};
    out.insert(out.end(), magic, magic + 4);
    size_t si_pos = out.size();
    out.resize(out.size() + 4 + 34, 0);
    Md5 md5;
    if (n * C > 0) {
        // int16 little-endian interleaved is exactly what the MD5 is defined over
        const uint16_t probe = 1;
        if (*(const uint8_t*)&probe == 1) {
            md5.update((const uint8_t*)pcm, (size_t)(n * C) * 2);
        } else {
            for (int64_t i = 0; i < n * C; ++i) { uint8_t b[2] = {(uint8_t)(pcm[i] & 0xff), (uint8_t)((uint16_t)pcm[i] >> 8)}; md5.update(b, 2); }
        }
    }
    std::vector<int32_t> chan, res;
    std::vector<uint32_t> zz;
    size_t min_frame = ~(size_t)0, max_frame = 0;
    static const int kRates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    int sr_code = 0;
    for (int i = 1; i < 12; ++i) if (kRates[i] == fs) sr_code = i;
    uint64_t frame_no = 0;
    for (int64_t pos = 0; pos < n; pos += kBlock, ++frame_no) {
        int bs = (int)((n - pos) < kBlock ? (n - pos) : kBlock);
        size_t start = out.size();
        BitWriter bw(out);
        bw.put(0x7ffc, 15);
        bw.put(0, 1);                                       // fixed block size stream
        int bs_code = bs == kBlock ? 12 : (bs <= 256 ? 6 : 7);
        bw.put((uint32_t)bs_code, 4);
        bw.put((uint32_t)sr_code, 4);
        bw.put((uint32_t)(C - 1), 4);                       // independent channels
        bw.put(4, 3);                                       // 16 bits per sample
        bw.put(0, 1);
        put_utf8(bw, frame_no);
        if (bs_code == 6) bw.put((uint32_t)(bs - 1), 8);
        else if (bs_code == 7) bw.put((uint32_t)(bs - 1), 16);
        out.push_back(crc8(out.data() + start, out.size() - start));
        chan.resize((size_t)bs);
        for (int c = 0; c < C; ++c) {
            for (int i = 0; i < bs; ++i) chan[(size_t)i] = pcm[(pos + i) * C + c];
            encode_subframe(bw, chan.data(), bs, 16, res, zz);
        }
        bw.align();
        uint16_t c16 = crc16(out.data() + start, out.size() - start);
        out.push_back((uint8_t)(c16 >> 8));
        out.push_back((uint8_t)(c16 & 0xff));
        size_t fl = out.size() - start;
        if (fl < min_frame) min_frame = fl;
        if (fl > max_frame) max_frame = fl;
    }
    if (frame_no == 0) min_frame = 0;
    uint8_t* h = out.data() + si_pos;
    h[0] = 0x80;                                            // last metadata block, type 0 (STREAMINFO)
    h[1] = 0; h[2] = 0; h[3] = 34;
    uint8_t* s = h + 4;
    s[0] = (uint8_t)(kBlock >> 8); s[1] = (uint8_t)(kBlock & 0xff);
    s[2] = s[0]; s[3] = s[1];
    s[4] = (uint8_t)(min_frame >> 16); s[5] = (uint8_t)(min_frame >> 8); s[6] = (uint8_t)min_frame;
    s[7] = (uint8_t)(max_frame >> 16); s[8] = (uint8_t)(max_frame >> 8); s[9] = (uint8_t)max_frame;
    s[10] = (uint8_t)(fs >> 12);
    s[11] = (uint8_t)(fs >> 4);
    s[12] = (uint8_t)(((fs & 0x0f) << 4) | ((C - 1) << 1) | 0);      // bps - 1 = 15 = 0b01111: high bit 0
    s[13] = (uint8_t)((15 << 4) | (int)((n >> 32) & 0x0f));
    s[14] = (uint8_t)(n >> 24); s[15] = (uint8_t)(n >> 16); s[16] = (uint8_t)(n >> 8); s[17] = (uint8_t)n;
    md5.finish(s + 18);
    return AIO_OK;
}

void encode_wav(const int16_t* pcm, int64_t n, int C, int fs, std::vector<uint8_t>& out) {
    uint32_t data = (uint32_t)(n * C * 2);
    out.resize(44 + (size_t)data);
    uint8_t* h = out.data();
    auto w32 = [&](int o, uint32_t v) { h[o] = (uint8_t)v; h[o + 1] = (uint8_t)(v >> 8); h[o + 2] = (uint8_t)(v >> 16); h[o + 3] = (uint8_t)(v >> 24); };
    auto w16 = [&](int o, uint32_t v) { h[o] = (uint8_t)v; h[o + 1] = (uint8_t)(v >> 8); };
    memcpy(h, "RIFF", 4); w32(4, 36 + data); memcpy(h + 8, "WAVEfmt ", 8); w32(16, 16);
    w16(20, 1); w16(22, (uint32_t)C); w32(24, (uint32_t)fs); w32(28, (uint32_t)(fs * C * 2)); w16(32, (uint32_t)(C * 2)); w16(34, 16);
    memcpy(h + 36, "data", 4); w32(40, data);
    for (int64_t i = 0; i < n * C; ++i) { h[44 + 2 * i] = (uint8_t)(pcm[i] & 0xff); h[45 + 2 * i] = (uint8_t)((uint16_t)pcm[i] >> 8); }
}

int probe_memory(const uint8_t* d, int64_t n, aio_info* info) {
    if (!d || !info) return AIO_ERR_INVALID;
    if (n < 4) return AIO_ERR_FORMAT;
    if (memcmp(d, "RIFF", 4) == 0) {
        WavInfo w;
        int rc = parse_wav(d, n, &w);
        if (rc) return rc;
        info->format = AIO_FMT_WAV;
        info->sample_rate = w.sample_rate;
        info->channels = w.channels;
        info->bits_per_sample = w.bps;
        info->n_samples = w.channels > 0 && w.bps >= 8 ? (int64_t)(w.data_len / (size_t)(w.channels * (w.bps / 8))) : 0;
        return AIO_OK;
    }
    StreamInfo si;
    int rc = parse_flac_header(d, n, &si);
    if (rc) return rc;
    info->format = AIO_FMT_FLAC;
    info->sample_rate = si.sample_rate;
    info->channels = si.channels;
    info->bits_per_sample = si.bps;
    info->n_samples = si.total > 0 ? si.total : -1;
    return AIO_OK;
}

}  // namespace

// ============================================================================= C-ABI
extern "C" {

int aio_probe_memory(const uint8_t* data, int64_t n_bytes, aio_info* info) { return probe_memory(data, n_bytes, info); }

int aio_probe_file(const char* path, aio_info* info) {
    if (!path || !info) return AIO_ERR_INVALID;
    std::vector<uint8_t> buf;
    // headers live in the first KBs, except a WAV whose data length we clamp to the file size:
    // read 64 KB, then fix up n_samples from the real file size for WAV.
    int rc = read_whole_file(path, buf, 1 << 16);
    if (rc) return rc;
    if (buf.size() >= 4 && memcmp(buf.data(), "RIFF", 4) == 0) {
        WavInfo w;
        rc = parse_wav(buf.data(), (int64_t)buf.size(), &w);
        if (rc) return rc;
        FILE* f = fopen(path, "rb");
        if (!f) return AIO_ERR_IO;
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fclose(f);
        size_t declared = le32(buf.data() + w.data_off - 4);
        size_t avail = (size_t)sz - w.data_off;
        size_t len = declared < avail ? declared : avail;
        info->format = AIO_FMT_WAV;
        info->sample_rate = w.sample_rate;
        info->channels = w.channels;
        info->bits_per_sample = w.bps;
        info->n_samples = w.channels > 0 && w.bps >= 8 ? (int64_t)(len / (size_t)(w.channels * (w.bps / 8))) : 0;
        return AIO_OK;
    }
    rc = probe_memory(buf.data(), (int64_t)buf.size(), info);
    if (rc == AIO_ERR_FORMAT && buf.size() == (1 << 16)) {   // metadata (pictures) longer than the probe window
        rc = read_whole_file(path, buf);
        if (rc) return rc;
        rc = probe_memory(buf.data(), (int64_t)buf.size(), info);
    }
    return rc;
}

int aio_decode_memory(const uint8_t* data, int64_t n_bytes, int16_t* out, int64_t capacity, int64_t* n_decoded,
                      int check_md5) {
    if (!data || !n_decoded || capacity < 0 || (!out && capacity > 0)) return AIO_ERR_INVALID;
    if (n_bytes < 4) return AIO_ERR_FORMAT;
    if (memcmp(data, "RIFF", 4) == 0) return decode_wav(data, n_bytes, out, capacity, n_decoded);
    return decode_flac(data, n_bytes, out, capacity, n_decoded, check_md5);
}

int aio_decode_file(const char* path, int16_t* out, int64_t capacity, int64_t* n_decoded, int check_md5) {
    if (!path) return AIO_ERR_INVALID;
    std::vector<uint8_t> buf;
    int rc = read_whole_file(path, buf);
    if (rc) return rc;
    return aio_decode_memory(buf.data(), (int64_t)buf.size(), out, capacity, n_decoded, check_md5);
}

int aio_probe_files(const char* const* paths, int32_t n, int32_t n_threads, aio_info* info, int32_t* status) {
    if (n < 0 || (n > 0 && (!paths || !info))) return AIO_ERR_INVALID;
    std::vector<int32_t> st((size_t)n, 0);
    parallel_for(n, n_threads, [&](int32_t i) { st[(size_t)i] = aio_probe_file(paths[i], &info[i]); });
    int first = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (status) status[i] = st[(size_t)i];
        if (!first && st[(size_t)i]) first = st[(size_t)i];
    }
    return first;
}

int aio_decode_files(const char* const* paths, int32_t n, int32_t n_threads, int16_t* out, const int64_t* offsets,
                     const int64_t* lengths, int check_md5, int32_t* status) {
    if (n < 0 || (n > 0 && (!paths || !out || !offsets || !lengths))) return AIO_ERR_INVALID;
    std::vector<int32_t> st((size_t)n, 0);
    parallel_for(n, n_threads, [&](int32_t i) {
        std::vector<uint8_t> buf;
        int rc = read_whole_file(paths[i], buf);
        if (!rc) {
            aio_info inf;
            rc = probe_memory(buf.data(), (int64_t)buf.size(), &inf);
            if (!rc && inf.channels != 1) rc = AIO_ERR_UNSUPPORTED;
        }
        if (!rc) {
            int64_t got = 0;
            rc = aio_decode_memory(buf.data(), (int64_t)buf.size(), out + offsets[i], lengths[i], &got, check_md5);
            if (!rc && got != lengths[i]) rc = AIO_ERR_FORMAT;
        }
        st[(size_t)i] = rc;
    });
    int first = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (status) status[i] = st[(size_t)i];
        if (!first && st[(size_t)i]) first = st[(size_t)i];
    }
    return first;
}

int aio_flac_layout(const uint8_t* data, int64_t n_bytes, aio_flac_layout_t* out) {
    if (!data || !out) return AIO_ERR_INVALID;
    StreamInfo si;
    int rc = parse_flac_header(data, n_bytes, &si);
    if (rc) return rc;
    out->n_samples = si.total;
    out->first_frame = (int32_t)si.first_frame;
    out->min_block = si.min_block;
    out->max_block = si.max_block;
    // a stream may declare min == max block size and still use the variable-blocking frame headers (sync 0xFFF9: the
    // header then numbers samples, not frames): report it as not fixed-block so that the device decoder declines it
    if (si.first_frame + 2 <= n_bytes && data[si.first_frame] == 0xFF && (data[si.first_frame + 1] & 0xFE) == 0xF8 &&
        (data[si.first_frame + 1] & 1))
        out->min_block = 0;
    out->sample_rate = si.sample_rate;
    out->channels = si.channels;
    out->bits_per_sample = si.bps;
    return AIO_OK;
}

int aio_flac_layouts(const uint8_t* buf, const int64_t* offsets, const int64_t* sizes, int32_t n, int32_t n_threads,
                     aio_flac_layout_t* out, int32_t* status) {
    if (n < 0 || (n > 0 && (!buf || !offsets || !sizes || !out))) return AIO_ERR_INVALID;
    std::vector<int32_t> st((size_t)n, 0);
    parallel_for(n, n_threads, [&](int32_t i) { st[(size_t)i] = aio_flac_layout(buf + offsets[i], sizes[i], &out[i]); });
    int first = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (status) status[i] = st[(size_t)i];
        if (!first && st[(size_t)i]) first = st[(size_t)i];
    }
    return first;
}

int aio_file_sizes(const char* const* paths, int32_t n, int64_t* sizes) {
    if (n < 0 || (n > 0 && (!paths || !sizes))) return AIO_ERR_INVALID;
    int first = 0;
    for (int32_t i = 0; i < n; ++i) {
        FILE* f = fopen(paths[i], "rb");
        long sz = -1;
        if (f) { if (fseek(f, 0, SEEK_END) == 0) sz = ftell(f); fclose(f); }
        sizes[i] = sz;
        if (sz < 0 && !first) first = AIO_ERR_IO;
    }
    return first;
}

int aio_read_files(const char* const* paths, int32_t n, int32_t n_threads, uint8_t* buf, const int64_t* offsets,
                   const int64_t* sizes, int32_t* status) {
    if (n < 0 || (n > 0 && (!paths || !buf || !offsets || !sizes))) return AIO_ERR_INVALID;
    std::vector<int32_t> st((size_t)n, 0);
    parallel_for(n, n_threads, [&](int32_t i) {
        FILE* f = fopen(paths[i], "rb");
        if (!f) { st[(size_t)i] = AIO_ERR_IO; return; }
        size_t want = (size_t)sizes[i];
        size_t got = want ? fread(buf + offsets[i], 1, want, f) : 0;
        fclose(f);
        st[(size_t)i] = got == want ? AIO_OK : AIO_ERR_IO;
    });
    int first = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (status) status[i] = st[(size_t)i];
        if (!first && st[(size_t)i]) first = st[(size_t)i];
    }
    return first;
}

int64_t aio_flac_bound(int64_t n_samples, int32_t channels) {
    if (n_samples < 0 || channels < 1) return -1;
    int64_t frames = (n_samples + kBlock - 1) / kBlock;
    return 42 + frames * (16 + 2 + (int64_t)channels * 2) + n_samples * channels * 2 + 64;
}

int aio_encode_flac(const int16_t* pcm, int64_t n_samples, int32_t channels, int32_t sample_rate, uint8_t* out,
                    int64_t capacity, int64_t* n_bytes) {
    if ((!pcm && n_samples > 0) || !out || !n_bytes) return AIO_ERR_INVALID;
    std::vector<uint8_t> buf;
    int rc = encode_flac(pcm, n_samples, channels, sample_rate, buf);
    if (rc) return rc;
    if ((int64_t)buf.size() > capacity) return AIO_ERR_CAPACITY;
    memcpy(out, buf.data(), buf.size());
    *n_bytes = (int64_t)buf.size();
    return AIO_OK;
}

int aio_write_file(const char* path, const int16_t* pcm, int64_t n_samples, int32_t channels, int32_t sample_rate,
                   int32_t format) {
    if (!path || (!pcm && n_samples > 0) || n_samples < 0 || channels < 1) return AIO_ERR_INVALID;
    std::vector<uint8_t> buf;
    if (format == AIO_FMT_FLAC) {
        int rc = encode_flac(pcm, n_samples, channels, sample_rate, buf);
        if (rc) return rc;
    } else if (format == AIO_FMT_WAV) {
        if (n_samples * channels * 2 > 0xffffff00ll) return AIO_ERR_UNSUPPORTED;
        encode_wav(pcm, n_samples, channels, sample_rate, buf);
    } else {
        return AIO_ERR_INVALID;
    }
    FILE* f = fopen(path, "wb");
    if (!f) return AIO_ERR_IO;
    size_t put = buf.empty() ? 0 : fwrite(buf.data(), 1, buf.size(), f);
    int bad = fclose(f);
    return (put == buf.size() && !bad) ? AIO_OK : AIO_ERR_IO;
}

int aio_write_files(const char* const* paths, int32_t n, int32_t n_threads, const int16_t* pcm, const int64_t* offsets,
                    const int64_t* lengths, int32_t sample_rate, int32_t format, int32_t* status) {
    if (n < 0 || (n > 0 && (!paths || !pcm || !offsets || !lengths))) return AIO_ERR_INVALID;
    std::vector<int32_t> st((size_t)n, 0);
    parallel_for(n, n_threads, [&](int32_t i) {
        st[(size_t)i] = aio_write_file(paths[i], pcm + offsets[i], lengths[i], 1, sample_rate, format);
    });
    int first = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (status) status[i] = st[(size_t)i];
        if (!first && st[(size_t)i]) first = st[(size_t)i];
    }
    return first;
}

const char* aio_strerror(int code) {
    switch (code) {
        case AIO_OK: return "ok";
        case AIO_ERR_INVALID: return "invalid argument";
        case AIO_ERR_IO: return "file open / read / write failed";
        case AIO_ERR_FORMAT: return "not a FLAC or WAV stream, or corrupt (sync, CRC, MD5, truncated)";
        case AIO_ERR_UNSUPPORTED: return "unsupported stream (more than 16 bits per sample, non-PCM WAV, or not mono)";
        case AIO_ERR_CAPACITY: return "output buffer too small";
        default: return "unknown error";
    }
}

}  // extern "C"
