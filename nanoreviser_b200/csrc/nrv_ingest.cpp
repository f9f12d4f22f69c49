// Host-side fast5 ingest at GPU rate (row A1 of section 8(a)).
//
// Replaces, for the common case, the h5py + Python event loop of the reference's get_read_data
// (nanorevutils/nanorev_fast5_handeler.py:39-150): N single-read Albacore fast5 files are parsed by a pool of host
// threads (own minimal HDF5 reader: there is no libhdf5 in the image), their event tables are collapsed to bases
// (:84-118: move 0 skip / 1 one base / 2 two bases / else one base), and the reads that succeed are packed straight into
// the CSR batch of include/nrv.h (signal = raw[a0:], starts relative to a0, ASCII bases, event mean / stdv, last_dur).
//
// HDF5 subset: superblock v0, object header v1 (+ continuation blocks), symbol-table groups
// (B-tree v1 / SNOD / local heap), dataspace v1/v2, datatypes int / float / fixed string / compound v1 / vlen string
// (global heap, for the `version` attribute), layout v3 contiguous + rank-1 chunked (B-tree v1 type 1) with the deflate
// filter (zlib) or ONT's VBZ filter (id 32020: zig-zag delta + streamvbyte + zstd; libzstd.so.1 is loaded at run time),
// attribute v1.  Legacy event tables (Albacore <= 0.0, :65-72: float64 `start` in seconds, rescaled with the raw read's
// `start_time`) are decoded like the reference does; multi-read containers (/read_<id>/{Raw/Signal, Analyses/...}, which the
// reference itself cannot open) yield one read per member.  Anything else outside the subset is NOT guessed at: the file gets
// status NRV_INGEST_UNSUPPORTED and the caller routes it through the Python reader (nanoreviser_b200/fast5.py), which
// follows the reference branch by branch.  Errors the reference raises map to the other status codes.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <dlfcn.h>
#include <math.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "nrv.h"

namespace {

constexpr uint64_t UNDEF = 0xFFFFFFFFFFFFFFFFull;

struct Fail {
    int code;
};

struct Buf {
    const uint8_t* p = nullptr;
    size_t n = 0;
    void need(uint64_t off, uint64_t len) const {
        if (off > n || len > n - off) throw Fail{NRV_INGEST_CORRUPT};
    }
    uint8_t u8(uint64_t o) const { need(o, 1); return p[o]; }
    uint16_t u16(uint64_t o) const { need(o, 2); uint16_t v; memcpy(&v, p + o, 2); return v; }
    uint32_t u32(uint64_t o) const { need(o, 4); uint32_t v; memcpy(&v, p + o, 4); return v; }
    uint64_t u64(uint64_t o) const { need(o, 8); uint64_t v; memcpy(&v, p + o, 8); return v; }
};

struct Msg {
    uint16_t type;
    uint64_t off;      // offset of the message data in the file
    uint16_t size;
};

inline uint64_t pad8(uint64_t n) { return (n + 7) & ~(uint64_t)7; }

struct H5 {
    Buf b;

    std::vector<Msg> object_header(uint64_t addr) const {
        if (b.u8(addr) != 1) throw Fail{NRV_INGEST_UNSUPPORTED};          // object header v1 only
        const uint16_t nmsg = b.u16(addr + 2);
        const uint32_t hsize = b.u32(addr + 8);
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, hsize}};
        std::vector<Msg> out;
        for (size_t bi = 0; bi < blocks.size() && out.size() < nmsg; ++bi) {
            uint64_t p = blocks[bi].first;
            const uint64_t end = p + blocks[bi].second;
            while (p + 8 <= end && out.size() < nmsg) {
                const uint16_t t = b.u16(p), sz = b.u16(p + 2);
                b.need(p + 8, sz);
                if (t == 0x0010) {
                    if (blocks.size() > 64) throw Fail{NRV_INGEST_CORRUPT};
                    blocks.push_back({b.u64(p + 8), b.u64(p + 16)});
                }
                out.push_back({t, p + 8, sz});
                p += 8 + sz;
            }
        }
        return out;
    }

    // children of a symbol-table group, in B-tree order
    void group_members(const std::vector<Msg>& msgs, std::vector<std::pair<std::string, uint64_t>>& out) const {
        for (const Msg& m : msgs) {
            if (m.type != 0x0011) continue;
            const uint64_t btree = b.u64(m.off), heap = b.u64(m.off + 8);
            b.need(heap, 32);
            if (memcmp(b.p + heap, "HEAP", 4)) throw Fail{NRV_INGEST_CORRUPT};
            const uint64_t dseg = b.u64(heap + 24);
            size_t budget = 1 << 16;                      // nodes visited: a cyclic tree fails instead of exploding
            walk_group(btree, dseg, out, 0, budget);
        }
    }
    void walk_group(uint64_t node, uint64_t dseg, std::vector<std::pair<std::string, uint64_t>>& out, int depth, size_t& budget) const {
        if (depth > 16 || budget-- == 0) throw Fail{NRV_INGEST_CORRUPT};
        b.need(node, 8);
        if (!memcmp(b.p + node, "TREE", 4)) {
            const uint16_t used = b.u16(node + 6);
            uint64_t p = node + 24;
            for (int i = 0; i < used; ++i, p += 16) walk_group(b.u64(p + 8), dseg, out, depth + 1, budget);
        } else if (!memcmp(b.p + node, "SNOD", 4)) {
            const uint16_t n = b.u16(node + 6);
            for (int i = 0; i < n; ++i) {
                const uint64_t e = node + 8 + 40ull * i;
                const uint64_t noff = b.u64(e), oaddr = b.u64(e + 8);
                const uint64_t s = dseg + noff;
                b.need(s, 1);
                const void* z = memchr(b.p + s, 0, b.n - s);
                if (!z) throw Fail{NRV_INGEST_CORRUPT};
                out.push_back({std::string((const char*)b.p + s, (const char*)z), oaddr});
            }
        } else {
            throw Fail{NRV_INGEST_CORRUPT};
        }
    }
    bool child(uint64_t group_addr, const std::string& name, uint64_t* addr) const {
        std::vector<std::pair<std::string, uint64_t>> mem;
        group_members(object_header(group_addr), mem);
        for (auto& kv : mem)
            if (kv.first == name) { *addr = kv.second; return true; }
        return false;
    }
    bool resolve(uint64_t root, const std::string& path, uint64_t* addr) const {
        uint64_t cur = root;
        size_t i = 0;
        while (i < path.size()) {
            while (i < path.size() && path[i] == '/') ++i;
            size_t j = i;
            while (j < path.size() && path[j] != '/') ++j;
            if (j > i && !child(cur, path.substr(i, j - i), &cur)) return false;
            i = j;
        }
        *addr = cur;
        return true;
    }

    // global-heap object (vlen data)
    std::string vlen(uint64_t ref_off) const {
        const uint32_t length = b.u32(ref_off);
        const uint64_t gaddr = b.u64(ref_off + 4);
        const uint32_t idx = b.u32(ref_off + 12);
        b.need(gaddr, 16);
        if (memcmp(b.p + gaddr, "GCOL", 4)) throw Fail{NRV_INGEST_CORRUPT};
        const uint64_t csize = b.u64(gaddr + 8);
        uint64_t p = gaddr + 16;
        const uint64_t end = gaddr + csize;
        while (p + 16 <= end) {
            const uint16_t oidx = b.u16(p);
            const uint64_t osize = b.u64(p + 8);
            if (oidx == idx) { b.need(p + 16, length); return std::string((const char*)b.p + p + 16, length); }
            if (oidx == 0) break;
            p += 16 + pad8(osize);
        }
        throw Fail{NRV_INGEST_CORRUPT};
    }
};

struct Member { int cls = -1; uint32_t size = 0, offset = 0; bool is_signed = false; bool present = false; };

// datatype message at `off`: returns its length in bytes; fills class / size (and compound members by name)
struct Dtype {
    int cls = -1;
    uint32_t size = 0;
    bool is_signed = false;
    uint64_t nbytes = 0;
};

Dtype parse_dtype(const Buf& b, uint64_t off, std::vector<std::pair<std::string, Member>>* members, int depth = 0) {
    if (depth > 4) throw Fail{NRV_INGEST_CORRUPT};
    Dtype d;
    const uint8_t cv = b.u8(off);
    d.cls = cv & 0x0F;
    const int ver = cv >> 4;
    const uint8_t b1 = b.u8(off + 1), b2 = b.u8(off + 2);
    d.size = b.u32(off + 4);
    uint64_t p = off + 8;
    switch (d.cls) {
        case 0: if (b1 & 1) throw Fail{NRV_INGEST_UNSUPPORTED}; d.is_signed = (b1 & 0x08) != 0; d.nbytes = 12; break;
        case 1: if (b1 & 1) throw Fail{NRV_INGEST_UNSUPPORTED}; d.nbytes = 20; break;
        case 3: d.nbytes = 8; break;
        case 6: {
            if (ver != 1) throw Fail{NRV_INGEST_UNSUPPORTED};
            const int nmemb = b1 | (b2 << 8);
            for (int i = 0; i < nmemb; ++i) {
                b.need(p, 1);
                const void* z = memchr(b.p + p, 0, b.n - p);
                if (!z) throw Fail{NRV_INGEST_CORRUPT};
                std::string name((const char*)b.p + p, (const char*)z);
                p += pad8(name.size() + 1);
                Member m;
                m.offset = b.u32(p);
                if (b.u8(p + 4) != 0) throw Fail{NRV_INGEST_UNSUPPORTED};    // array members
                p += 4 + 1 + 3 + 4 + 4 + 16;
                const Dtype sub = parse_dtype(b, p, nullptr, depth + 1);
                p += sub.nbytes;
                m.cls = sub.cls; m.size = sub.size; m.is_signed = sub.is_signed; m.present = true;
                if (members) members->push_back({name, m});
            }
            d.nbytes = p - off;
            break;
        }
        case 9: {
            const Dtype base = parse_dtype(b, p, nullptr, depth + 1);
            d.nbytes = 8 + base.nbytes;
            break;
        }
        default: throw Fail{NRV_INGEST_UNSUPPORTED};
    }
    return d;
}

void parse_dataspace(const Buf& b, uint64_t off, std::vector<uint64_t>& dims) {
    const uint8_t ver = b.u8(off), rank = b.u8(off + 1);
    uint64_t p;
    if (ver == 1) p = off + 8; else if (ver == 2) p = off + 4; else throw Fail{NRV_INGEST_UNSUPPORTED};
    dims.clear();
    for (int i = 0; i < rank; ++i) dims.push_back(b.u64(p + 8ull * i));
}

// "version" attribute of /Analyses/<group>: true if LooseVersion(v) <= LooseVersion('0.0') (fast5_handeler.py:65-68)
bool version_le_zero(const std::string& v) {
    std::vector<long> parts;
    std::string cur;
    auto flush = [&]() {
        if (cur.empty()) return;
        char* e = nullptr;
        const long x = strtol(cur.c_str(), &e, 10);
        parts.push_back((*e == 0) ? x : 1);            // any alphabetic tag sorts above '0.0'
        cur.clear();
    };
    for (char c : v) {
        if (c == '.' || c == '-') flush(); else if (c != 0) cur.push_back(c);
    }
    flush();
    while (parts.size() < 2) parts.push_back(0);
    const std::vector<long> zero{0, 0};
    return !std::lexicographical_compare(zero.begin(), zero.end(), parts.begin(), parts.end());
}

// ---- VBZ (HDF5 filter 32020, nanoporetech/vbz_compression; the reference bundles the plugin binary under nanorevutils/utils/lib).
// chunk = u32 decompressed byte count, then -- inside a zstd frame unless cd_values[3] == 0 -- streamvbyte: ceil(n / 4) control
// bytes (2 bits per value, low bits first: byte count - 1), then the values' little-endian bytes; with cd_values[2] the values are
// zig-zag coded deltas.  cd_values = {version, integer size, zig-zag flag, zstd level}; versions 0 and 1 do not differ for 2-byte
// integers.  Pinned against outputs of the plugin binary itself (tests/golden/make_variant_fixtures.py, tests/test_ingest_variants.py).
struct Zstd {
    unsigned long long (*frame_size)(const void*, size_t) = nullptr;
    size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    bool ok = false;
};
const Zstd& zstd_lib() {
    static Zstd z;
    static std::once_flag once;
    std::call_once(once, []() {
        void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        z.frame_size = reinterpret_cast<unsigned long long (*)(const void*, size_t)>(dlsym(h, "ZSTD_getFrameContentSize"));
        z.decompress = reinterpret_cast<size_t (*)(void*, size_t, const void*, size_t)>(dlsym(h, "ZSTD_decompress"));
        z.is_error = reinterpret_cast<unsigned (*)(size_t)>(dlsym(h, "ZSTD_isError"));
        z.ok = z.frame_size && z.decompress && z.is_error;
    });
    return z;
}

struct Filter { int id = 0; uint32_t ncd = 0; uint32_t cd[8] = {0, 0, 0, 0, 0, 0, 0, 0}; };

// one VBZ chunk of int16 samples -> dst (at most max_samples); returns the number of samples produced
size_t vbz_decode_i16(const uint8_t* src, size_t n, const Filter& f, int16_t* dst, size_t max_samples, std::vector<uint8_t>& tmp) {
    if (f.ncd < 3 || f.cd[1] != 2) throw Fail{NRV_INGEST_UNSUPPORTED};      // 2-byte integers only (fast5 signals)
    const bool zigzag = f.cd[2] != 0;
    const uint32_t zlevel = f.ncd > 3 ? f.cd[3] : 1;
    if (n < 4) throw Fail{NRV_INGEST_CORRUPT};
    uint32_t out_bytes;
    memcpy(&out_bytes, src, 4);
    const size_t count = out_bytes / 2;
    const uint8_t* body = src + 4;
    size_t blen = n - 4;
    if (zlevel) {
        const Zstd& z = zstd_lib();
        if (!z.ok) throw Fail{NRV_INGEST_UNSUPPORTED};                        // no libzstd on this host: the Python reader reports it
        const unsigned long long fs = z.frame_size(body, blen);
        if (fs > (5ull * count + 16)) throw Fail{NRV_INGEST_CORRUPT};         // also catches ZSTD_CONTENTSIZE_UNKNOWN / _ERROR
        tmp.resize((size_t)fs + 8);
        const size_t r = z.decompress(tmp.data(), (size_t)fs, body, blen);
        if (z.is_error(r) || r != fs) throw Fail{NRV_INGEST_CORRUPT};
        body = tmp.data(); blen = (size_t)fs;
    }
    const size_t nctl = (count + 3) / 4;
    if (blen < nctl) throw Fail{NRV_INGEST_CORRUPT};
    const uint8_t* data = body + nctl;
    size_t avail = blen - nctl, pos = 0;
    uint32_t prev = 0;
    const size_t n_out = std::min(count, max_samples);
    for (size_t i = 0; i < n_out; ++i) {
        const int len = ((body[i >> 2] >> ((i & 3) * 2)) & 3) + 1;
        if (pos + (size_t)len > avail) throw Fail{NRV_INGEST_CORRUPT};
        uint32_t v = 0;
        for (int k = 0; k < len; ++k) v |= (uint32_t)data[pos + k] << (8 * k);
        pos += (size_t)len;
        if (zigzag) { prev += (v >> 1) ^ (0u - (v & 1)); v = prev; }
        dst[i] = (int16_t)(uint16_t)v;
    }
    return n_out;
}

struct ReadOut {
    int status = NRV_INGEST_OPEN_FAILED;
    int64_t a0 = 0;
    int32_t last_dur = 0;
    std::vector<int32_t> starts;
    std::vector<uint8_t> bases;
    std::vector<float> ev_mean, ev_std;
    std::vector<int16_t> signal;       // raw[a0:]
    std::vector<uint8_t> qual;         // basecaller Phred scores of the bases (empty when the Fastq dataset is absent / does not line up)
    std::string name;                  // multi-read containers: the member's name (read_<id>); empty for single-read files
};

double load_real(const uint8_t* p, const Member& m) {
    if (m.cls == 1 && m.size == 4) { float v; memcpy(&v, p, 4); return v; }
    if (m.cls == 1 && m.size == 8) { double v; memcpy(&v, p, 8); return v; }
    throw Fail{NRV_INGEST_UNSUPPORTED};
}
int64_t load_int(const uint8_t* p, const Member& m) {
    if (m.cls != 0) throw Fail{NRV_INGEST_UNSUPPORTED};
    switch (m.size) {
        case 1: return m.is_signed ? (int64_t)(int8_t)p[0] : (int64_t)p[0];
        case 2: { uint16_t v; memcpy(&v, p, 2); return m.is_signed ? (int64_t)(int16_t)v : (int64_t)v; }
        case 4: { uint32_t v; memcpy(&v, p, 4); return m.is_signed ? (int64_t)(int32_t)v : (int64_t)v; }
        case 8: { uint64_t v; memcpy(&v, p, 8); return (int64_t)v; }
    }
    throw Fail{NRV_INGEST_UNSUPPORTED};
}

// integer attribute `name` of the object at `addr` (attribute v1); false if absent
bool int_attr(const H5& h, uint64_t addr, const char* name, int64_t* out) {
    for (const Msg& m : h.object_header(addr)) {
        if (m.type != 0x000C) continue;
        if (h.b.u8(m.off) != 1) throw Fail{NRV_INGEST_UNSUPPORTED};
        const uint16_t nsz = h.b.u16(m.off + 2), dtsz = h.b.u16(m.off + 4), dssz = h.b.u16(m.off + 6);
        uint64_t p = m.off + 8;
        h.b.need(p, nsz);
        const std::string nm((const char*)h.b.p + p, strnlen((const char*)h.b.p + p, nsz));
        p += pad8(nsz);
        if (nm != name) continue;
        const Dtype dt = parse_dtype(h.b, p, nullptr);
        p += pad8(dtsz) + pad8(dssz);
        if (dt.cls != 0) throw Fail{NRV_INGEST_UNSUPPORTED};
        Member mm; mm.cls = 0; mm.size = dt.size; mm.is_signed = dt.is_signed;
        h.b.need(p, dt.size);
        *out = load_int(h.b.p + p, mm);
        return true;
    }
    return false;
}

// One read.  `top` = the object that holds Analyses/ (the file's root group, or a read_<id> member of a multi-read container);
// `multi`: the signal is top/Raw/Signal and the raw attributes sit on top/Raw, instead of /Raw/Reads/<first>/Signal (:132-133).
void parse_read(const H5& h, uint64_t top, bool multi, const std::string& group, const std::string& subgroup, ReadOut& out,
                std::vector<uint8_t>& scratch) {
    out = ReadOut();
    try {
        const uint64_t root = top;
        // ---- raw read (:130-135): first child of /Raw/Reads; located first because a legacy table needs its start_time ----
        uint64_t raddr = UNDEF, saddr = UNDEF, rattr = UNDEF;
        int raw_fail = 0;
        try {
            if (multi) {
                if (!h.resolve(root, "Raw", &rattr) || !h.child(rattr, "Signal", &saddr)) throw Fail{NRV_INGEST_NO_SIGNAL};
            } else {
                std::vector<std::pair<std::string, uint64_t>> reads;
                if (!h.resolve(root, "/Raw/Reads", &raddr)) throw Fail{NRV_INGEST_NO_SIGNAL};
                h.group_members(h.object_header(raddr), reads);
                if (reads.empty()) throw Fail{NRV_INGEST_NO_SIGNAL};
                if (reads.size() > 1) std::sort(reads.begin(), reads.end());               // h5py iterates members by name
                rattr = reads[0].second;
                if (!h.child(rattr, "Signal", &saddr)) throw Fail{NRV_INGEST_NO_SIGNAL};
            }
        } catch (Fail f) {
            raw_fail = f.code == NRV_INGEST_UNSUPPORTED ? NRV_INGEST_UNSUPPORTED : NRV_INGEST_NO_SIGNAL;   // reported after the event stage, as the reference does
        }
        // ---- events (:63-77) ----
        uint64_t gaddr, eaddr;
        int stage = NRV_INGEST_NO_EVENTS;
        bool legacy = false;
        int64_t start_time = 0;
        try {
            if (!h.resolve(root, "Analyses/" + group, &gaddr)) throw Fail{NRV_INGEST_NO_EVENTS};
            std::string version = "0.0";
            for (const Msg& m : h.object_header(gaddr)) {
                if (m.type != 0x000C) continue;
                if (h.b.u8(m.off) != 1) throw Fail{NRV_INGEST_UNSUPPORTED};
                const uint16_t nsz = h.b.u16(m.off + 2), dtsz = h.b.u16(m.off + 4), dssz = h.b.u16(m.off + 6);
                uint64_t p = m.off + 8;
                h.b.need(p, nsz);
                const std::string name((const char*)h.b.p + p, strnlen((const char*)h.b.p + p, nsz));
                p += pad8(nsz);
                if (name != "version") continue;
                const Dtype dt = parse_dtype(h.b, p, nullptr);
                p += pad8(dtsz) + pad8(dssz);
                if (dt.cls == 9) version = h.vlen(p);
                else if (dt.cls == 3) { h.b.need(p, dt.size); version.assign((const char*)h.b.p + p, strnlen((const char*)h.b.p + p, dt.size)); }
                else throw Fail{NRV_INGEST_UNSUPPORTED};
            }
            if (version_le_zero(version)) {                                                // legacy rescaling branch (:65-73)
                legacy = true;
                // raw_attrs = attrs of the first raw read; a missing read or start_time is the reference's "No events" error (:76-78)
                if (rattr == UNDEF || !int_attr(h, rattr, "start_time", &start_time)) throw Fail{NRV_INGEST_NO_EVENTS};
            }
            if (!h.resolve(gaddr, subgroup + "/Events", &eaddr)) throw Fail{NRV_INGEST_NO_EVENTS};
        } catch (Fail f) {
            throw Fail{f.code == NRV_INGEST_UNSUPPORTED ? NRV_INGEST_UNSUPPORTED : stage};
        }
        std::vector<std::pair<std::string, Member>> members;
        std::vector<uint64_t> dims;
        uint32_t itemsize = 0;
        uint64_t ev_addr = UNDEF;
        for (const Msg& m : h.object_header(eaddr)) {
            if (m.type == 0x0001) parse_dataspace(h.b, m.off, dims);
            else if (m.type == 0x0003) { const Dtype dt = parse_dtype(h.b, m.off, &members); if (dt.cls != 6) throw Fail{NRV_INGEST_NO_EVENTS}; itemsize = dt.size; }
            else if (m.type == 0x0008) {
                if (h.b.u8(m.off) != 3 || h.b.u8(m.off + 1) != 1) throw Fail{NRV_INGEST_UNSUPPORTED};   // contiguous only
                ev_addr = h.b.u64(m.off + 2);
            } else if (m.type == 0x000B) throw Fail{NRV_INGEST_UNSUPPORTED};
        }
        if (dims.size() != 1 || !itemsize || ev_addr == UNDEF) throw Fail{NRV_INGEST_NO_EVENTS};
        Member m_mean, m_start, m_stdv, m_state, m_move;
        for (auto& kv : members) {
            if (kv.first == "mean") m_mean = kv.second;
            else if (kv.first == "start") m_start = kv.second;
            else if (kv.first == "stdv") m_stdv = kv.second;
            else if (kv.first == "model_state") m_state = kv.second;
            else if (kv.first == "move") m_move = kv.second;
        }
        if (!m_mean.present || !m_start.present || !m_stdv.present || !m_state.present || !m_move.present) throw Fail{NRV_INGEST_NO_EVENTS};
        // legacy tables hold `start` in seconds as float64: start * 4000 - start_time, written back into the float64 column and
        // truncated by int() (:71,:93).  Any other combination (integer starts in a legacy file, float32 / float starts in a
        // current one) goes to the Python reader, which has numpy's casting rules
        const bool fstart = m_start.cls == 1 && m_start.size == 8;
        if (legacy ? !fstart : m_start.cls != 0) throw Fail{NRV_INGEST_UNSUPPORTED};
        if (m_state.cls != 3 || m_state.size < 3) throw Fail{NRV_INGEST_UNSUPPORTED};
        // every member the collapse reads must lie inside one record (a corrupt compound type must not send the loops below
        // past the file buffer)
        for (const Member* mm : {&m_mean, &m_start, &m_stdv, &m_state, &m_move})
            if ((uint64_t)mm->offset + mm->size > itemsize) throw Fail{NRV_INGEST_CORRUPT};
        const uint64_t E = dims[0];
        if (E > h.b.n / itemsize) throw Fail{NRV_INGEST_CORRUPT};                          // E * itemsize cannot wrap below
        h.b.need(ev_addr, E * itemsize);
        const uint8_t* ev = h.b.p + ev_addr;
        // ---- collapse events to bases (:84-118), forward order ----
        size_t nb = 0;
        for (uint64_t i = 0; i < E; ++i) {
            const int64_t mv = load_int(ev + i * itemsize + m_move.offset, m_move);
            nb += (mv == 0) ? 0 : (mv == 2 ? 2 : 1);
        }
        if (nb < 2) throw Fail{NRV_INGEST_TOO_SHORT};                                      // np.diff / start[-2] fail (:120-128)
        std::vector<int64_t> start(nb);
        out.bases.resize(nb); out.ev_mean.resize(nb); out.ev_std.resize(nb);
        size_t k = 0;
        for (uint64_t i = 0; i < E; ++i) {
            const uint8_t* r = ev + i * itemsize;
            const int64_t mv = load_int(r + m_move.offset, m_move);
            if (mv == 0) continue;
            int64_t st;
            if (legacy) {
                const double sv = load_real(r + m_start.offset, m_start) * 4000.0 - (double)start_time;
                if (!(fabs(sv) < 9.0e15)) throw Fail{NRV_INGEST_CORRUPT};
                st = (int64_t)sv;                                                              // int(): truncation toward zero
            } else {
                st = load_int(r + m_start.offset, m_start);
            }
            const float mean = (float)load_real(r + m_mean.offset, m_mean), sd = (float)load_real(r + m_stdv.offset, m_stdv);
            const uint8_t* ms = r + m_state.offset;
            if (mv == 2) {
                start[k] = st; out.bases[k] = ms[1]; out.ev_mean[k] = mean; out.ev_std[k] = sd; ++k;
                start[k] = st + 2; out.bases[k] = ms[2]; out.ev_mean[k] = mean; out.ev_std[k] = sd; ++k;
            } else {
                start[k] = st; out.bases[k] = ms[2]; out.ev_mean[k] = mean; out.ev_std[k] = sd; ++k;
            }
        }
        const int last_dur = (start[nb - 1] - start[nb - 2] < 5) ? 3 : 5;                  // :121-126
        // ---- raw signal (:130-135) ----
        if (raw_fail) throw Fail{raw_fail};
        std::vector<uint64_t> sdims;
        Dtype sdt;
        int layout_cls = -1;
        uint64_t lay_off = 0;
        std::vector<Filter> filters;
        for (const Msg& m : h.object_header(saddr)) {
            if (m.type == 0x0001) parse_dataspace(h.b, m.off, sdims);
            else if (m.type == 0x0003) sdt = parse_dtype(h.b, m.off, nullptr);
            else if (m.type == 0x0008) { if (h.b.u8(m.off) != 3) throw Fail{NRV_INGEST_UNSUPPORTED}; layout_cls = h.b.u8(m.off + 1); lay_off = m.off; }
            else if (m.type == 0x000B) {
                if (h.b.u8(m.off) != 1) throw Fail{NRV_INGEST_UNSUPPORTED};
                const int nf = h.b.u8(m.off + 1);
                uint64_t p = m.off + 8;
                for (int i = 0; i < nf; ++i) {
                    const uint16_t fid = h.b.u16(p), name_len = h.b.u16(p + 2), ncd = h.b.u16(p + 6);
                    Filter f; f.id = fid; f.ncd = std::min<uint32_t>(ncd, 8);
                    for (uint32_t c = 0; c < f.ncd; ++c) f.cd[c] = h.b.u32(p + 8 + pad8(name_len) + 4ull * c);
                    p += 8 + pad8(name_len) + 4ull * ncd + ((ncd & 1) ? 4 : 0);
                    filters.push_back(f);
                }
            }
        }
        if (sdims.size() != 1 || sdt.cls != 0 || sdt.size != 2 || !sdt.is_signed) throw Fail{NRV_INGEST_UNSUPPORTED};
        const uint64_t S = sdims[0];
        if (S > ((uint64_t)1 << 34) || S / 1100 > h.b.n) throw Fail{NRV_INGEST_CORRUPT};      // deflate expands at most ~1032x
        if ((int64_t)S < start[nb - 1] + last_dur) throw Fail{NRV_INGEST_SIGNAL_SHORT};   // :142-143
        const int64_t a0 = start[0];
        if (a0 < 0 || (uint64_t)a0 > S) throw Fail{NRV_INGEST_SIGNAL_SHORT};
        std::vector<int16_t> sig(S, 0);
        if (layout_cls == 1) {
            const uint64_t addr = h.b.u64(lay_off + 2);
            if (addr != UNDEF) { h.b.need(addr, S * 2); memcpy(sig.data(), h.b.p + addr, S * 2); }
            if (!filters.empty()) throw Fail{NRV_INGEST_UNSUPPORTED};
        } else if (layout_cls == 2) {
            const int ndims = h.b.u8(lay_off + 2);
            if (ndims != 2) throw Fail{NRV_INGEST_UNSUPPORTED};                           // rank-1 datasets only
            const uint64_t btree = h.b.u64(lay_off + 3);
            const uint32_t cdim = h.b.u32(lay_off + 11), esize = h.b.u32(lay_off + 15);
            if (esize != 2) throw Fail{NRV_INGEST_UNSUPPORTED};
            if (cdim == 0 || cdim > (1u << 28)) throw Fail{NRV_INGEST_CORRUPT};
            for (const Filter& f : filters) if (f.id != 1 && f.id != 32020) throw Fail{NRV_INGEST_UNSUPPORTED};   // deflate or VBZ
            if (filters.size() > 1) throw Fail{NRV_INGEST_UNSUPPORTED};
            if (btree != UNDEF) {
                // chunk B-tree v1 (node type 1): keys {u32 size, u32 filter mask, ndims x u64 offsets}, children = chunk addresses
                // bounded walk: a corrupt (cyclic) tree cannot loop or grow the stack without limit
                std::vector<std::pair<uint64_t, int>> stack{{btree, 0}};
                size_t visited = 0;
                while (!stack.empty()) {
                    const uint64_t node = stack.back().first;
                    const int depth = stack.back().second;
                    stack.pop_back();
                    if (depth > 16 || ++visited > 65536) throw Fail{NRV_INGEST_CORRUPT};
                    h.b.need(node, 24);
                    if (memcmp(h.b.p + node, "TREE", 4) || h.b.u8(node + 4) != 1) throw Fail{NRV_INGEST_CORRUPT};
                    const int level = h.b.u8(node + 5), used = h.b.u16(node + 6);
                    const uint64_t keysz = 8 + 8ull * ndims;
                    uint64_t p = node + 24;
                    for (int i = 0; i < used; ++i, p += keysz + 8) {
                        const uint32_t csize = h.b.u32(p), fmask = h.b.u32(p + 4);
                        const uint64_t off0 = h.b.u64(p + 8), childaddr = h.b.u64(p + keysz);
                        if (level > 0) { stack.push_back({childaddr, depth + 1}); continue; }
                        if (off0 >= S) continue;
                        h.b.need(childaddr, csize);
                        const uint64_t room = (S - off0) * 2;
                        if (!filters.empty() && !(fmask & 1) && filters[0].id == 32020) {
                            vbz_decode_i16(h.b.p + childaddr, csize, filters[0], sig.data() + off0, (size_t)(S - off0), scratch);
                        } else if (!filters.empty() && !(fmask & 1)) {
                            // the chunk may be larger than the dataset and the inflated payload shorter than the chunk
                            scratch.resize((size_t)cdim * 2);
                            z_stream zs;
                            memset(&zs, 0, sizeof(zs));
                            if (inflateInit(&zs) != Z_OK) throw Fail{NRV_INGEST_CORRUPT};
                            zs.next_in = const_cast<Bytef*>(h.b.p + childaddr); zs.avail_in = csize;
                            zs.next_out = scratch.data(); zs.avail_out = (uInt)scratch.size();
                            const int rc = inflate(&zs, Z_FINISH);
                            const uint64_t produced = zs.total_out;
                            inflateEnd(&zs);
                            if (rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) throw Fail{NRV_INGEST_CORRUPT};
                            memcpy(reinterpret_cast<uint8_t*>(sig.data()) + off0 * 2, scratch.data(), (size_t)std::min<uint64_t>(produced, room));
                        } else {
                            memcpy(reinterpret_cast<uint8_t*>(sig.data()) + off0 * 2, h.b.p + childaddr, (size_t)std::min<uint64_t>(csize, room));
                        }
                    }
                }
            }
        } else {
            throw Fail{NRV_INGEST_UNSUPPORTED};
        }
        // ---- basecaller qualities for -F fastq (best effort, never fails the read): the event-collapsed call is the Fastq sequence
        //      without its first and last two bases (bases == Fastq_seq[2:-2]); qualities are Phred = character - 33, capped at 93 ----
        try {
            uint64_t qaddr;
            if (h.resolve(gaddr, subgroup + "/Fastq", &qaddr)) {
                Dtype qdt; uint64_t qdata = UNDEF; std::vector<uint64_t> qdims; bool filtered = false;
                for (const Msg& m : h.object_header(qaddr)) {
                    if (m.type == 0x0001) parse_dataspace(h.b, m.off, qdims);
                    else if (m.type == 0x0003) qdt = parse_dtype(h.b, m.off, nullptr);
                    else if (m.type == 0x0008) {
                        if (h.b.u8(m.off) == 3 && h.b.u8(m.off + 1) == 1) qdata = h.b.u64(m.off + 2);            // contiguous
                        else if (h.b.u8(m.off) == 3 && h.b.u8(m.off + 1) == 0) qdata = m.off + 4;                 // compact: u16 size, data
                    } else if (m.type == 0x000B) filtered = true;
                }
                if (qdt.cls == 3 && qdims.empty() && qdata != UNDEF && !filtered) {
                    h.b.need(qdata, qdt.size);
                    const char* t = (const char*)h.b.p + qdata;
                    const size_t n = strnlen(t, qdt.size);
                    // lines: @name \n sequence \n + \n qualities
                    const char* l1 = (const char*)memchr(t, '\n', n);
                    const char* l2 = l1 ? (const char*)memchr(l1 + 1, '\n', n - (size_t)(l1 + 1 - t)) : nullptr;
                    const char* l3 = l2 ? (const char*)memchr(l2 + 1, '\n', n - (size_t)(l2 + 1 - t)) : nullptr;
                    if (l3) {
                        const char* seq = l1 + 1; const size_t nseq = (size_t)(l2 - seq);
                        const char* ql = l3 + 1;
                        const char* l4 = (const char*)memchr(ql, '\n', n - (size_t)(ql - t));
                        const size_t nq = l4 ? (size_t)(l4 - ql) : n - (size_t)(ql - t);
                        if (nseq == nq && nseq == nb + 4 && !memcmp(seq + 2, out.bases.data(), nb)) {
                            out.qual.resize(nb);
                            for (size_t i = 0; i < nb; ++i) {
                                const int q = (int)(uint8_t)ql[i + 2] - 33;
                                out.qual[i] = (uint8_t)std::min(std::max(q, 0), 93);
                            }
                        }
                    }
                }
            }
        } catch (...) { out.qual.clear(); }
        // ---- per-read outputs in the C-ABI layout ----
        out.a0 = a0;
        out.last_dur = last_dur;
        out.starts.resize(nb);
        for (size_t i = 0; i < nb; ++i) {
            const int64_t rel = start[i] - a0;
            if (rel < 0 || rel > INT32_MAX) throw Fail{NRV_INGEST_UNSUPPORTED};
            out.starts[i] = (int32_t)rel;
        }
        out.signal.assign(sig.begin() + a0, sig.end());
        out.status = NRV_INGEST_OK;
    } catch (Fail f) {
        const int code = f.code;
        out = ReadOut();
        out.status = code;
    } catch (...) {
        out = ReadOut();
        out.status = NRV_INGEST_CORRUPT;
    }
}

// One file: a single-read fast5 (one ReadOut) or a multi-read container (one ReadOut per read_<id> member, in name order).
// file status: the read's status for a single-read file; for a container NRV_INGEST_OK if at least one member decoded, else the
// first member's failure.
void read_file(const char* path, const std::string& group, const std::string& subgroup, std::vector<ReadOut>& outs, int* file_status,
               std::vector<uint8_t>& filebuf, std::vector<uint8_t>& scratch) {
    outs.clear();
    *file_status = NRV_INGEST_OPEN_FAILED;
    // ---- whole file into memory (single-read fast5: 0.1 - 1 MB) ----
    FILE* fp = fopen(path, "rb");
    if (!fp) return;
    fseek(fp, 0, SEEK_END);
    const long fsz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (fsz < 96) { fclose(fp); return; }
    filebuf.resize((size_t)fsz);
    const size_t got = fread(filebuf.data(), 1, (size_t)fsz, fp);
    fclose(fp);
    if (got != (size_t)fsz) return;
    H5 h;
    h.b.p = filebuf.data(); h.b.n = filebuf.size();
    static const uint8_t SIG[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (memcmp(h.b.p, SIG, 8)) return;                                                      // "Error opening file" (:59-61)
    std::vector<std::pair<std::string, uint64_t>> top;
    uint64_t root = 0;
    try {
        if (h.b.u8(8) != 0 || h.b.u8(13) != 8 || h.b.u8(14) != 8) throw Fail{NRV_INGEST_UNSUPPORTED};
        root = h.b.u64(64);
        h.group_members(h.object_header(root), top);
    } catch (Fail f) {
        *file_status = f.code;
        return;
    } catch (...) {
        *file_status = NRV_INGEST_CORRUPT;
        return;
    }
    bool has_raw = false, has_read = false;
    for (auto& kv : top) { has_raw |= kv.first == "Raw" || kv.first == "Analyses"; has_read |= kv.first.compare(0, 5, "read_") == 0; }
    if (has_raw || !has_read) {
        outs.resize(1);
        parse_read(h, root, false, group, subgroup, outs[0], scratch);
        *file_status = outs[0].status;
        if (outs[0].status != NRV_INGEST_OK) outs.clear();
        return;
    }
    std::sort(top.begin(), top.end());
    int first_fail = NRV_INGEST_NO_EVENTS;
    bool any_fail = false;
    for (auto& kv : top) {
        if (kv.first.compare(0, 5, "read_") != 0) continue;
        ReadOut r;
        parse_read(h, kv.second, true, group, subgroup, r, scratch);
        if (r.status == NRV_INGEST_OK) { r.name = kv.first; outs.push_back(std::move(r)); }
        else if (!any_fail) { any_fail = true; first_fail = r.status; }
    }
    *file_status = outs.empty() ? first_fail : NRV_INGEST_OK;
}

}  // namespace

struct nrv_ingest {
    std::vector<int32_t> file_status;      // [n_files]
    std::vector<int64_t> read_file;        // [n_reads] index of the file each packed read came from
    std::vector<int64_t> a0;               // [n_reads]
    std::vector<int16_t> signal;
    std::vector<int64_t> sig_off, base_off;
    std::vector<int32_t> starts, last_dur;
    std::vector<uint8_t> bases;
    std::vector<float> ev_mean, ev_std;
    std::vector<uint8_t> qual;             // [n_bases] basecaller Phred scores, or empty when any packed read has none
    std::vector<std::string> names;        // [n_reads] member name inside a multi-read container, "" for a single-read file
    std::vector<const char*> name_ptrs;
};

extern "C" {

int nrv_ingest_fast5(const char* const* paths, int64_t n_files, const char* group, const char* subgroup, int n_threads,
                     nrv_ingest** out) {
    if (!out || n_files < 0 || (n_files > 0 && !paths)) return NRV_E_INVALID;
    const std::string g = group ? group : "Basecall_1D_000", sg = subgroup ? subgroup : "BaseCalled_template";
    std::vector<std::vector<ReadOut>> per((size_t)n_files);
    std::vector<int> fstat((size_t)n_files, NRV_INGEST_OPEN_FAILED);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n_files));
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        std::vector<uint8_t> filebuf, scratch;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_files) break;
            read_file(paths[i], g, sg, per[(size_t)i], &fstat[(size_t)i], filebuf, scratch);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();

    nrv_ingest* r = new nrv_ingest();
    r->file_status.resize((size_t)n_files);
    r->sig_off.push_back(0); r->base_off.push_back(0);
    int64_t ns = 0, nb = 0, nr = 0;
    bool all_qual = true;
    for (int64_t i = 0; i < n_files; ++i) {
        r->file_status[(size_t)i] = fstat[(size_t)i];
        for (const ReadOut& p : per[(size_t)i]) {
            ns += (int64_t)p.signal.size(); nb += (int64_t)p.starts.size(); ++nr;
            if (p.qual.size() != p.starts.size()) all_qual = false;
        }
    }
    all_qual = all_qual && nr > 0;
    r->signal.resize((size_t)ns); r->starts.resize((size_t)nb); r->bases.resize((size_t)nb);
    r->ev_mean.resize((size_t)nb); r->ev_std.resize((size_t)nb);
    r->last_dur.reserve((size_t)nr); r->a0.reserve((size_t)nr); r->read_file.reserve((size_t)nr); r->names.reserve((size_t)nr);
    if (all_qual) r->qual.resize((size_t)nb);
    int64_t so = 0, bo = 0;
    for (int64_t i = 0; i < n_files; ++i) {
        for (ReadOut& p : per[(size_t)i]) {
            if (all_qual) memcpy(r->qual.data() + bo, p.qual.data(), p.qual.size());
            memcpy(r->signal.data() + so, p.signal.data(), p.signal.size() * 2);
            memcpy(r->starts.data() + bo, p.starts.data(), p.starts.size() * 4);
            memcpy(r->bases.data() + bo, p.bases.data(), p.bases.size());
            memcpy(r->ev_mean.data() + bo, p.ev_mean.data(), p.ev_mean.size() * 4);
            memcpy(r->ev_std.data() + bo, p.ev_std.data(), p.ev_std.size() * 4);
            so += (int64_t)p.signal.size(); bo += (int64_t)p.starts.size();
            r->sig_off.push_back(so); r->base_off.push_back(bo);
            r->last_dur.push_back(p.last_dur); r->a0.push_back(p.a0); r->read_file.push_back(i);
            r->names.push_back(p.name);
            ReadOut().signal.swap(p.signal);           // release per-file copies as we go
        }
    }
    r->name_ptrs.resize(r->names.size());
    for (size_t i = 0; i < r->names.size(); ++i) r->name_ptrs[i] = r->names[i].c_str();
    *out = r;
    return NRV_OK;
}

int nrv_ingest_read_names(const nrv_ingest* r, const char* const** names) {
    if (!r || !names) return NRV_E_INVALID;
    *names = r->name_ptrs.data();
    return NRV_OK;
}

int nrv_ingest_view(const nrv_ingest* r, nrv_batch* batch, const int32_t** file_status, const int64_t** read_file, const int64_t** a0) {
    if (!r || !batch) return NRV_E_INVALID;
    batch->n_reads = (int64_t)r->last_dur.size();
    batch->signal = r->signal.data(); batch->sig_off = r->sig_off.data();
    batch->starts = r->starts.data(); batch->base_off = r->base_off.data();
    batch->bases = r->bases.data(); batch->ev_mean = r->ev_mean.data(); batch->ev_std = r->ev_std.data();
    batch->last_dur = r->last_dur.data();
    batch->qual = r->qual.empty() ? nullptr : r->qual.data();   // all-or-nothing: NULL unless every packed read has its qualities
    if (file_status) *file_status = r->file_status.data();
    if (read_file) *read_file = r->read_file.data();
    if (a0) *a0 = r->a0.data();
    return NRV_OK;
}

void nrv_ingest_free(nrv_ingest* r) { delete r; }

}  // extern "C"
