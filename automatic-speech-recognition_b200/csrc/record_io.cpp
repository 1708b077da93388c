// TFRecord / tf.train.Example writer and reader (include/asr_record_io.h) that serialises the
// front-end's flat float32 output buffer without going through Python objects.
//
// Replaces create_tfrecords (/root/reference/create_tfrecord.py:43-97) and provides the parser side
// of tfrecord_data_loader.py:24-52 for TF-free consumers and for the round-trip tests.  Written from
// the protobuf wire-format and TFRecord framing definitions; no TensorFlow code.
#include "../../include/asr_record_io.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <nmmintrin.h>
#endif

namespace {

// ----------------------------------------------------------------------------- CRC-32C
struct Crc32cTable {
    uint32_t t[8][256];
    Crc32cTable() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82f63b78u : c >> 1;   // reflected Castagnoli polynomial
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
    }
};
const Crc32cTable kT;

uint32_t crc32c_sw(uint32_t c, const uint8_t* p, size_t n) {
    while (n >= 8) {
        uint32_t lo, hi;
        memcpy(&lo, p, 4);
        memcpy(&hi, p + 4, 4);
        lo ^= c;
        c = kT.t[7][lo & 0xff] ^ kT.t[6][(lo >> 8) & 0xff] ^ kT.t[5][(lo >> 16) & 0xff] ^ kT.t[4][lo >> 24] ^
            kT.t[3][hi & 0xff] ^ kT.t[2][(hi >> 8) & 0xff] ^ kT.t[1][(hi >> 16) & 0xff] ^ kT.t[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) c = (c >> 8) ^ kT.t[0][(c ^ *p++) & 0xff];
    return c;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) uint32_t crc32c_hw(uint32_t c, const uint8_t* p, size_t n) {
    uint64_t c64 = c;
    while (n >= 8) {
        uint64_t v;
        memcpy(&v, p, 8);
        c64 = _mm_crc32_u64(c64, v);
        p += 8;
        n -= 8;
    }
    c = (uint32_t)c64;
    while (n--) c = _mm_crc32_u8(c, *p++);
    return c;
}
const bool kHaveHw = __builtin_cpu_supports("sse4.2");
#else
const bool kHaveHw = false;
uint32_t crc32c_hw(uint32_t c, const uint8_t* p, size_t n) { return crc32c_sw(c, p, n); }
#endif

inline uint32_t crc32c(const uint8_t* p, size_t n) {
    uint32_t c = 0xffffffffu;
    c = kHaveHw ? crc32c_hw(c, p, n) : crc32c_sw(c, p, n);
    return c ^ 0xffffffffu;
}
inline uint32_t mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

// ----------------------------------------------------------------------------- protobuf wire helpers
inline int varint_size(uint64_t v) {
    int n = 1;
    while (v >= 0x80) { v >>= 7; ++n; }
    return n;
}
inline uint8_t* put_varint(uint8_t* p, uint64_t v) {
    while (v >= 0x80) { *p++ = (uint8_t)(v | 0x80); v >>= 7; }
    *p++ = (uint8_t)v;
    return p;
}
inline int64_t int64_payload(const int64_t* v, int64_t n) {
    int64_t s = 0;
    for (int64_t i = 0; i < n; ++i) s += varint_size((uint64_t)v[i]);
    return s;
}

struct ExampleSizes {
    int64_t feat_list, feat_feature, feat_entry;
    int64_t shape_payload, shape_list, shape_feature, shape_entry;
    int64_t token_payload, token_list, token_feature, token_entry;
    int64_t features, total;
};

inline int64_t len_delimited(int64_t body) { return 1 + varint_size((uint64_t)body) + body; }

ExampleSizes example_sizes(int64_t n_feat, const int64_t* shape, int n_shape, const int64_t* token, int64_t n_token) {
    ExampleSizes z;
    z.feat_list = n_feat ? len_delimited(4 * n_feat) : 0;
    z.feat_feature = len_delimited(z.feat_list);
    z.feat_entry = len_delimited(4) + len_delimited(z.feat_feature);          // key "feat"
    z.shape_payload = int64_payload(shape, n_shape);
    z.shape_list = n_shape ? len_delimited(z.shape_payload) : 0;
    z.shape_feature = len_delimited(z.shape_list);
    z.shape_entry = len_delimited(5) + len_delimited(z.shape_feature);        // key "shape"
    z.token_payload = int64_payload(token, n_token);
    z.token_list = n_token ? len_delimited(z.token_payload) : 0;
    z.token_feature = len_delimited(z.token_list);
    z.token_entry = len_delimited(5) + len_delimited(z.token_feature);        // key "token"
    z.features = len_delimited(z.feat_entry) + len_delimited(z.shape_entry) + len_delimited(z.token_entry);
    z.total = len_delimited(z.features);
    return z;
}

uint8_t* put_int64_entry(uint8_t* p, const char* key, int key_len, const int64_t* v, int64_t n, int64_t entry,
                         int64_t feature, int64_t list, int64_t payload) {
    *p++ = 0x0a; p = put_varint(p, (uint64_t)entry);                          // Features.feature map entry
    *p++ = 0x0a; p = put_varint(p, (uint64_t)key_len); memcpy(p, key, (size_t)key_len); p += key_len;
    *p++ = 0x12; p = put_varint(p, (uint64_t)feature);                        // entry.value: Feature
    *p++ = 0x1a; p = put_varint(p, (uint64_t)list);                           // Feature.int64_list
    if (n) {
        *p++ = 0x0a; p = put_varint(p, (uint64_t)payload);                    // Int64List.value, packed
        for (int64_t i = 0; i < n; ++i) p = put_varint(p, (uint64_t)v[i]);
    }
    return p;
}

uint8_t* serialize_example(uint8_t* p, const ExampleSizes& z, const float* feat, int64_t n_feat, const int64_t* shape,
                           int n_shape, const int64_t* token, int64_t n_token) {
    *p++ = 0x0a; p = put_varint(p, (uint64_t)z.features);                     // Example.features
    *p++ = 0x0a; p = put_varint(p, (uint64_t)z.feat_entry);
    *p++ = 0x0a; *p++ = 4; memcpy(p, "feat", 4); p += 4;
    *p++ = 0x12; p = put_varint(p, (uint64_t)z.feat_feature);
    *p++ = 0x12; p = put_varint(p, (uint64_t)z.feat_list);                    // Feature.float_list
    if (n_feat) {
        *p++ = 0x0a; p = put_varint(p, (uint64_t)(4 * n_feat));               // FloatList.value, packed little-endian
        memcpy(p, feat, (size_t)(4 * n_feat));
        p += 4 * n_feat;
    }
    p = put_int64_entry(p, "shape", 5, shape, n_shape, z.shape_entry, z.shape_feature, z.shape_list, z.shape_payload);
    p = put_int64_entry(p, "token", 5, token, n_token, z.token_entry, z.token_feature, z.token_list, z.token_payload);
    return p;
}

// ----------------------------------------------------------------------------- protobuf reader
struct Span {
    const uint8_t* p;
    const uint8_t* e;
};
bool get_varint(Span& s, uint64_t* v) {
    uint64_t r = 0;
    for (int sh = 0; sh < 64 && s.p < s.e; sh += 7) {
        uint8_t b = *s.p++;
        r |= (uint64_t)(b & 0x7f) << sh;
        if (!(b & 0x80)) { *v = r; return true; }
    }
    return false;
}
bool get_len(Span& s, Span* out) {
    uint64_t n;
    if (!get_varint(s, &n) || n > (uint64_t)(s.e - s.p)) return false;
    out->p = s.p;
    out->e = s.p + n;
    s.p += n;
    return true;
}
bool skip_field(Span& s, int wt) {
    uint64_t v;
    Span t;
    switch (wt) {
        case 0: return get_varint(s, &v);
        case 1: if (s.e - s.p < 8) return false; s.p += 8; return true;
        case 2: return get_len(s, &t);
        case 5: if (s.e - s.p < 4) return false; s.p += 4; return true;
        default: return false;
    }
}

struct ParsedExample {
    std::vector<float> feat;
    std::vector<int64_t> shape, token;
};

bool parse_float_list(Span s, std::vector<float>* out) {
    while (s.p < s.e) {
        uint64_t tag;
        if (!get_varint(s, &tag)) return false;
        int fn = (int)(tag >> 3), wt = (int)(tag & 7);
        if (fn == 1 && wt == 2) {
            Span b;
            if (!get_len(s, &b) || ((b.e - b.p) & 3)) return false;
            size_t n = (size_t)(b.e - b.p) / 4, o = out->size();
            out->resize(o + n);
            memcpy(out->data() + o, b.p, n * 4);
        } else if (fn == 1 && wt == 5) {
            if (s.e - s.p < 4) return false;
            float f;
            memcpy(&f, s.p, 4);
            s.p += 4;
            out->push_back(f);
        } else if (!skip_field(s, wt)) {
            return false;
        }
    }
    return true;
}
bool parse_int64_list(Span s, std::vector<int64_t>* out) {
    while (s.p < s.e) {
        uint64_t tag, v;
        if (!get_varint(s, &tag)) return false;
        int fn = (int)(tag >> 3), wt = (int)(tag & 7);
        if (fn == 1 && wt == 2) {
            Span b;
            if (!get_len(s, &b)) return false;
            while (b.p < b.e) {
                if (!get_varint(b, &v)) return false;
                out->push_back((int64_t)v);
            }
        } else if (fn == 1 && wt == 0) {
            if (!get_varint(s, &v)) return false;
            out->push_back((int64_t)v);
        } else if (!skip_field(s, wt)) {
            return false;
        }
    }
    return true;
}

bool parse_example(const uint8_t* d, size_t n, ParsedExample* ex) {
    Span s{d, d + n};
    while (s.p < s.e) {
        uint64_t tag;
        if (!get_varint(s, &tag)) return false;
        if ((tag >> 3) == 1 && (tag & 7) == 2) {                              // Example.features
            Span fs;
            if (!get_len(s, &fs)) return false;
            while (fs.p < fs.e) {
                if (!get_varint(fs, &tag)) return false;
                if ((tag >> 3) != 1 || (tag & 7) != 2) { if (!skip_field(fs, (int)(tag & 7))) return false; continue; }
                Span entry;
                if (!get_len(fs, &entry)) return false;
                std::string key;
                Span value{nullptr, nullptr};
                while (entry.p < entry.e) {
                    if (!get_varint(entry, &tag)) return false;
                    Span t;
                    if ((tag & 7) != 2) { if (!skip_field(entry, (int)(tag & 7))) return false; continue; }
                    if (!get_len(entry, &t)) return false;
                    if ((tag >> 3) == 1) key.assign((const char*)t.p, (size_t)(t.e - t.p));
                    else if ((tag >> 3) == 2) value = t;
                }
                if (!value.p) continue;
                while (value.p < value.e) {                                     // Feature: oneof kind
                    if (!get_varint(value, &tag)) return false;
                    Span body;
                    if ((tag & 7) != 2) { if (!skip_field(value, (int)(tag & 7))) return false; continue; }
                    if (!get_len(value, &body)) return false;
                    int fn = (int)(tag >> 3);
                    if (fn == 2 && key == "feat") { if (!parse_float_list(body, &ex->feat)) return false; }
                    else if (fn == 3 && key == "shape") { if (!parse_int64_list(body, &ex->shape)) return false; }
                    else if (fn == 3 && key == "token") { if (!parse_int64_list(body, &ex->token)) return false; }
                }
            }
        } else if (!skip_field(s, (int)(tag & 7))) {
            return false;
        }
    }
    return true;
}

// walks the records of a file; fn(record bytes, index) returns false to stop with a format error
template <class F>
int64_t for_each_record(const char* path, F fn) {
    FILE* f = fopen(path, "rb");
    if (!f) return RIO_ERR_IO;
    std::vector<uint8_t> buf;
    int64_t count = 0;
    for (;;) {
        uint8_t hdr[12];
        size_t got = fread(hdr, 1, 12, f);
        if (got == 0) break;
        if (got != 12) { fclose(f); return RIO_ERR_FORMAT; }
        uint64_t len;
        uint32_t lcrc;
        memcpy(&len, hdr, 8);
        memcpy(&lcrc, hdr + 8, 4);
        if (mask_crc(crc32c(hdr, 8)) != lcrc || len > (1ull << 31)) { fclose(f); return RIO_ERR_FORMAT; }   // protobuf's own limit
        buf.resize((size_t)len + 4);
        if (fread(buf.data(), 1, (size_t)len + 4, f) != (size_t)len + 4) { fclose(f); return RIO_ERR_FORMAT; }
        uint32_t dcrc;
        memcpy(&dcrc, buf.data() + len, 4);
        if (mask_crc(crc32c(buf.data(), (size_t)len)) != dcrc) { fclose(f); return RIO_ERR_FORMAT; }
        if (!fn(buf.data(), (size_t)len, count)) { fclose(f); return RIO_ERR_FORMAT; }
        ++count;
    }
    fclose(f);
    return count;
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

int write_range(const char* path, int32_t lo, int32_t hi, const float* feats, const int64_t* feat_offsets,
                const int32_t* n_frames, int32_t feat_dim, int32_t planes, const int64_t* tokens,
                const int64_t* token_offsets, const int32_t* token_lens) {
    FILE* f = fopen(path, "wb");
    if (!f) return RIO_ERR_IO;
    std::vector<char> iobuf(1 << 20);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    std::vector<uint8_t> rec;
    bool ok = true;
    for (int32_t r = lo; r < hi && ok; ++r) {
        const int64_t shape[3] = {n_frames[r], feat_dim, planes};
        const int n_shape = planes > 0 ? 3 : 2;                               // (L, D) matrices when args.cmvn is false
        const int64_t n_feat = (int64_t)n_frames[r] * feat_dim * (planes > 0 ? planes : 1);
        const int64_t* tok = tokens ? tokens + token_offsets[r] : nullptr;
        const int64_t n_tok = tokens ? token_lens[r] : 0;
        ExampleSizes z = example_sizes(n_feat, shape, n_shape, tok, n_tok);
        rec.resize((size_t)z.total + 16);
        uint64_t len = (uint64_t)z.total;
        memcpy(rec.data(), &len, 8);
        uint32_t c = mask_crc(crc32c(rec.data(), 8));
        memcpy(rec.data() + 8, &c, 4);
        uint8_t* end = serialize_example(rec.data() + 12, z, feats + feat_offsets[r], n_feat, shape, n_shape, tok, n_tok);
        if (end != rec.data() + 12 + z.total) { ok = false; break; }
        c = mask_crc(crc32c(rec.data() + 12, (size_t)z.total));
        memcpy(end, &c, 4);
        ok = fwrite(rec.data(), 1, rec.size(), f) == rec.size();
    }
    if (fclose(f) != 0) ok = false;
    return ok ? RIO_OK : RIO_ERR_IO;
}

}  // namespace

extern "C" {

uint32_t rio_crc32c(const void* data, int64_t n_bytes) { return crc32c((const uint8_t*)data, (size_t)(n_bytes > 0 ? n_bytes : 0)); }
uint32_t rio_masked_crc32c(const void* data, int64_t n_bytes) { return mask_crc(rio_crc32c(data, n_bytes)); }

int64_t rio_example_size(int64_t n_feat, const int64_t* shape, int32_t n_shape, const int64_t* token, int64_t n_token) {
    if (n_feat < 0 || n_shape < 0 || n_token < 0 || (n_shape && !shape) || (n_token && !token)) return RIO_ERR_INVALID;
    return example_sizes(n_feat, shape, n_shape, token, n_token).total;
}

int rio_example_serialize(const float* feat, int64_t n_feat, const int64_t* shape, int32_t n_shape, const int64_t* token,
                          int64_t n_token, uint8_t* out, int64_t capacity, int64_t* n_bytes) {
    if (n_feat < 0 || n_shape < 0 || n_token < 0 || (n_feat && !feat) || (n_shape && !shape) || (n_token && !token) ||
        !out || !n_bytes)
        return RIO_ERR_INVALID;
    ExampleSizes z = example_sizes(n_feat, shape, n_shape, token, n_token);
    if (z.total > capacity) return RIO_ERR_CAPACITY;
    uint8_t* end = serialize_example(out, z, feat, n_feat, shape, n_shape, token, n_token);
    *n_bytes = (int64_t)(end - out);
    return *n_bytes == z.total ? RIO_OK : RIO_ERR_FORMAT;
}

int rio_write_tfrecord(const char* path, int32_t n, const float* feats, const int64_t* feat_offsets, const int32_t* n_frames,
                       int32_t feat_dim, int32_t planes, const int64_t* tokens, const int64_t* token_offsets,
                       const int32_t* token_lens) {
    if (!path || n < 0 || feat_dim < 1 || planes < 0 || (n > 0 && (!feats || !feat_offsets || !n_frames)) ||
        (tokens && (!token_offsets || !token_lens)))
        return RIO_ERR_INVALID;
    return write_range(path, 0, n, feats, feat_offsets, n_frames, feat_dim, planes, tokens, token_offsets, token_lens);
}

int rio_write_tfrecords(const char* const* paths, int32_t n_files, const int32_t* file_start, int32_t n_threads,
                        const float* feats, const int64_t* feat_offsets, const int32_t* n_frames, int32_t feat_dim,
                        int32_t planes, const int64_t* tokens, const int64_t* token_offsets, const int32_t* token_lens,
                        int32_t* status) {
    if (n_files < 0 || (n_files > 0 && (!paths || !file_start || !feats || !feat_offsets || !n_frames)) || feat_dim < 1 ||
        planes < 0 || (tokens && (!token_offsets || !token_lens)))
        return RIO_ERR_INVALID;
    std::vector<int32_t> st((size_t)n_files, 0);
    parallel_for(n_files, n_threads, [&](int32_t f) {
        st[(size_t)f] = write_range(paths[f], file_start[f], file_start[f + 1], feats, feat_offsets, n_frames, feat_dim,
                                    planes, tokens, token_offsets, token_lens);
    });
    int first = 0;
    for (int32_t f = 0; f < n_files; ++f) {
        if (status) status[f] = st[(size_t)f];
        if (!first && st[(size_t)f]) first = st[(size_t)f];
    }
    return first;
}

int64_t rio_index_tfrecord(const char* path, int64_t capacity, int64_t* n_feat, int64_t* shapes3, int64_t* n_token) {
    if (!path) return RIO_ERR_INVALID;
    const bool detail = n_feat || shapes3 || n_token;
    bool overflow = false;
    int64_t rc = for_each_record(path, [&](const uint8_t* d, size_t n, int64_t i) {
        if (!detail) return true;
        if (i >= capacity) { overflow = true; return true; }
        ParsedExample ex;
        if (!parse_example(d, n, &ex)) return false;
        if (n_feat) n_feat[i] = (int64_t)ex.feat.size();
        if (n_token) n_token[i] = (int64_t)ex.token.size();
        if (shapes3)
            for (int k = 0; k < 3; ++k) shapes3[3 * i + k] = k < (int)ex.shape.size() ? ex.shape[(size_t)k] : 0;
        return true;
    });
    if (rc >= 0 && overflow) return RIO_ERR_CAPACITY;
    return rc;
}

int rio_read_tfrecord(const char* path, int64_t n, float* feats, const int64_t* feat_offsets, int64_t* tokens,
                      const int64_t* token_offsets) {
    if (!path || n < 0 || (n > 0 && (!feats || !feat_offsets)) || (tokens && !token_offsets)) return RIO_ERR_INVALID;
    int64_t rc = for_each_record(path, [&](const uint8_t* d, size_t len, int64_t i) {
        if (i >= n) return false;
        ParsedExample ex;
        if (!parse_example(d, len, &ex)) return false;
        if (!ex.feat.empty()) memcpy(feats + feat_offsets[i], ex.feat.data(), ex.feat.size() * 4);
        if (tokens && !ex.token.empty()) memcpy(tokens + token_offsets[i], ex.token.data(), ex.token.size() * 8);
        return true;
    });
    if (rc < 0) return (int)rc;
    return rc == n ? RIO_OK : RIO_ERR_FORMAT;
}

const char* rio_strerror(int code) {
    switch (code) {
        case RIO_OK: return "ok";
        case RIO_ERR_INVALID: return "invalid argument";
        case RIO_ERR_IO: return "file open / read / write failed";
        case RIO_ERR_FORMAT: return "corrupt TFRecord (length, CRC-32C or protobuf)";
        case RIO_ERR_CAPACITY: return "output buffer too small";
        default: return "unknown error";
    }
}

}  // extern "C"
