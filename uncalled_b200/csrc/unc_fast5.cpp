// unc_fast5.cpp -- host-side fast5 input: what Fast5Reader + ReadBuffer(hdf5_tools::File&, raw_path, ch_path)
// do in the reference (src/fast5_reader.cpp:134-248, src/read_buffer.cpp:198-246), without libhdf5.
//
// The reference vendors libhdf5 1.8.21 and reads three things per read: the attributes of the read's Raw
// group and of its channel_id group (every attribute converted to TEXT by hdf5_tools' String_Reader and then
// parsed with atoi/atof, submods/fast5/include/fast5/hdf5_tools.hpp:1013-1141) and the int16 dataset `Signal`.
// This file reads exactly that subset of the HDF5 file format, as laid down in the HDF5 File Format
// Specification (versions 1.0-3.0): superblock v0-v3, object headers v1 and v2, old-style groups (symbol-table
// B-tree + local heap), compact new-style groups (link messages), attribute messages v1-v3 (fixed-point,
// floating-point, fixed and variable-length string; global heap), dataspaces v1/v2, data layouts v1-v3
// (compact, contiguous, chunked through the v1 B-tree) and the deflate / shuffle / fletcher32 filters.
// What it does not read is reported as an error, never guessed: dense link / attribute storage (fractal
// heaps), layout v4, the VBZ filter (id 32020; the reference's vendored libhdf5 cannot read it either without an
// external plugin).
//
// Signals are delivered as the raw int16 DAC values plus the calibration constants; the i16 -> u16
// reinterpretation and the float calibration of src/read_buffer.cpp:239-242 run on the GPU (unc_sample).
// unc_fast5_load decodes a range of reads with several host threads straight into a caller-provided
// (pinned) staging buffer.
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/unc_b200.h"

namespace {

struct H5Error : std::runtime_error { using std::runtime_error::runtime_error; };

const uint64_t UNDEF = ~0ull;

struct Msg { uint16_t type; uint8_t flags; const uint8_t *p; size_t n; };

struct DType {
    int cls = -1;            // 0 fixed-point, 1 floating-point, 3 string, 9 variable-length
    uint32_t size = 0;
    bool is_signed = false, big_endian = false;
    int str_pad = 0;         // 0 null-terminated, 1 null-padded, 2 space-padded
    bool vlen_string = false;
};

struct DSpace { int rank = 0; uint64_t dims[4] = {0, 0, 0, 0}; uint64_t count = 1; bool null_space = false; };

struct Filter { uint16_t id; std::vector<uint32_t> cd; };

struct Dataset {
    DType dt; DSpace ds;
    int layout = -1;         // 0 compact, 1 contiguous, 2 chunked
    const uint8_t *compact = nullptr; size_t compact_size = 0;
    uint64_t addr = UNDEF, size = 0;
    uint64_t chunk_elems = 0;
    int chunk_rank = 0;      // dimensionality stored in the layout message (rank + 1)
    std::vector<Filter> filters;
};

struct H5File {
    const uint8_t *d = nullptr;
    size_t size = 0;
    int fd = -1;
    unsigned O = 8, L = 8;
    uint64_t base = 0, root = UNDEF;

    ~H5File() {
        if (d) munmap((void *) d, size);
        if (fd >= 0) close(fd);
    }
    const uint8_t *at(uint64_t off, uint64_t n) const {
        if (off > size || n > size - off) throw H5Error("truncated or corrupt file (offset beyond end)");
        return d + off;
    }
    static uint64_t le(const uint8_t *p, unsigned n) {
        uint64_t v = 0;
        for (unsigned i = 0; i < n; i++) v |= (uint64_t) p[i] << (8 * i);
        return v;
    }
    uint64_t off_at(const uint8_t *p) const {       // an "offset"-sized field; all ones = undefined
        uint64_t v = le(p, O);
        if (O < 8 && v == ((1ull << (8 * O)) - 1)) return UNDEF;
        return v;
    }
    uint64_t len_at(const uint8_t *p) const { return le(p, L); }
    const uint8_t *abs(uint64_t a, uint64_t n) const {
        if (a == UNDEF) throw H5Error("undefined address");
        if (base > size || a > size - base) throw H5Error("truncated or corrupt file (address beyond end)");   // no wrap-around in base + a
        return at(base + a, n);
    }

    void open(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) throw H5Error(std::string("cannot open ") + path + ": " + strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < 16) throw H5Error(std::string(path) + ": not an HDF5 file");
        size = (size_t) st.st_size;
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { d = nullptr; throw H5Error(std::string("mmap failed: ") + strerror(errno)); }
        d = (const uint8_t *) m;
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        uint64_t sb = UNDEF;
        for (uint64_t o = 0; o + 8 <= size; o = o ? o * 2 : 512)      // superblock at 0, 512, 1024, ...
            if (!memcmp(d + o, sig, 8)) { sb = o; break; }
        if (sb == UNDEF) throw H5Error(std::string(path) + ": not an HDF5 file (no superblock signature)");
        const uint8_t *p = at(sb, 16);
        const unsigned ver = p[8];
        if (ver <= 1) {
            at(sb, 24);
            O = p[13]; L = p[14];
            if ((O != 4 && O != 8) || (L != 4 && L != 8)) throw H5Error("unsupported offset/length size");
            uint64_t q = sb + 24 + (ver == 1 ? 4 : 0);
            base = off_at(at(q, O));
            q += 4 * (uint64_t) O;                                        // base, free space, end of file, driver info
            const uint8_t *ste = at(q, 2 * O + 24);                       // root group symbol table entry
            root = off_at(ste + O);
        } else if (ver <= 3) {
            O = p[9]; L = p[10];
            if ((O != 4 && O != 8) || (L != 4 && L != 8)) throw H5Error("unsupported offset/length size");
            const uint8_t *q = at(sb + 12, 4 * (uint64_t) O);
            base = off_at(q);
            root = off_at(q + 3 * O);
        } else {
            throw H5Error("unsupported superblock version");
        }
        if (base == UNDEF) base = 0;
        if (ver <= 1 && base == 0 && sb != 0) base = sb;                  // user block: addresses are relative to it
    }

    // ---- object headers ------------------------------------------------------------------------------
    void parse_block_v1(uint64_t addr, uint64_t len, unsigned &left, std::vector<Msg> &out,
                        std::vector<std::pair<uint64_t, uint64_t>> &cont) const {
        const uint8_t *p = abs(addr, len);
        uint64_t q = 0;
        while (left && q + 8 <= len) {
            Msg m;
            m.type = (uint16_t) le(p + q, 2);
            m.n = (size_t) le(p + q + 2, 2);
            m.flags = p[q + 4];
            if (q + 8 + m.n > len) throw H5Error("corrupt object header (message overruns its block)");
            m.p = p + q + 8;
            q += 8 + m.n;
            left--;
            if (m.type == 0x10) {
                if (m.n < O + L) throw H5Error("corrupt continuation message");
                cont.push_back({off_at(m.p), len_at(m.p + O)});
            } else if (m.type != 0) {
                out.push_back(m);
            }
        }
    }
    void parse_block_v2(const uint8_t *p, uint64_t len, bool order, std::vector<Msg> &out,
                        std::vector<std::pair<uint64_t, uint64_t>> &cont) const {
        uint64_t q = 0;
        const unsigned hdr = order ? 6 : 4;
        while (q + hdr <= len) {
            Msg m;
            m.type = p[q];
            m.n = (size_t) le(p + q + 1, 2);
            m.flags = p[q + 3];
            if (q + hdr + m.n > len) break;                               // a gap smaller than a message header
            m.p = p + q + hdr;
            q += hdr + m.n;
            if (m.type == 0x10) {
                if (m.n < O + L) throw H5Error("corrupt continuation message");
                cont.push_back({off_at(m.p), len_at(m.p + O)});
            } else if (m.type != 0) {
                out.push_back(m);
            }
        }
    }
    std::vector<Msg> messages(uint64_t addr) const {
        std::vector<Msg> out;
        std::vector<std::pair<uint64_t, uint64_t>> cont;
        const uint8_t *p = abs(addr, 16);
        if (!memcmp(p, "OHDR", 4)) {
            if (p[4] != 2) throw H5Error("unsupported object header version");
            const uint8_t fl = p[5];
            uint64_t q = 6;
            if (fl & 0x20) q += 16;
            if (fl & 0x10) q += 4;
            const unsigned sz = 1u << (fl & 3);
            const uint64_t chunk0 = le(abs(addr + q, sz), sz);
            q += sz;
            parse_block_v2(abs(addr + q, chunk0), chunk0, fl & 4, out, cont);
            for (size_t i = 0; i < cont.size(); i++) {
                if (cont.size() > 4096) throw H5Error("corrupt object header (continuation loop)");
                const uint8_t *c = abs(cont[i].first, cont[i].second);
                if (cont[i].second < 8 || memcmp(c, "OCHK", 4)) throw H5Error("corrupt object header continuation");
                parse_block_v2(c + 4, cont[i].second - 8, fl & 4, out, cont);
            }
        } else {
            if (p[0] != 1) throw H5Error("unsupported object header version");
            unsigned left = (unsigned) le(p + 2, 2);
            const uint64_t hsize = le(p + 8, 4);
            parse_block_v1(addr + 16, hsize, left, out, cont);
            for (size_t i = 0; i < cont.size() && left; i++) {
                if (cont.size() > 4096) throw H5Error("corrupt object header (continuation loop)");
                parse_block_v1(cont[i].first, cont[i].second, left, out, cont);
            }
        }
        return out;
    }

    // ---- groups ----------------------------------------------------------------------------------------
    void walk_group_btree(uint64_t node, uint64_t heap_data, uint64_t heap_size, int depth,
                          std::vector<std::pair<std::string, uint64_t>> &out) const {
        if (depth > 32) throw H5Error("corrupt group B-tree");
        const uint8_t *p = abs(node, 8 + 2 * (uint64_t) O);
        if (memcmp(p, "TREE", 4) || p[4] != 0) throw H5Error("corrupt group B-tree node");
        const unsigned level = p[5], used = (unsigned) le(p + 6, 2);
        const uint8_t *e = abs(node + 8 + 2 * O, (uint64_t) used * (L + O) + L);
        for (unsigned i = 0; i < used; i++) {
            const uint64_t child = off_at(e + L + (uint64_t) i * (L + O));
            if (level > 0) { walk_group_btree(child, heap_data, heap_size, depth + 1, out); continue; }
            const uint8_t *s = abs(child, 8);
            if (memcmp(s, "SNOD", 4)) throw H5Error("corrupt symbol table node");
            const unsigned nsym = (unsigned) le(s + 6, 2);
            const uint64_t esz = 2 * (uint64_t) O + 24;
            const uint8_t *ent = abs(child + 8, nsym * esz);
            for (unsigned k = 0; k < nsym; k++) {
                const uint8_t *t = ent + k * esz;
                const uint64_t name_off = off_at(t), obj = off_at(t + O);
                const uint32_t cache = (uint32_t) le(t + 2 * O, 4);
                if (cache == 2 || obj == UNDEF) continue;                 // soft link
                if (name_off >= heap_size) throw H5Error("corrupt symbol table entry");
                const char *nm = (const char *) abs(heap_data + name_off, 1);
                const size_t maxn = (size_t) (heap_size - name_off);
                out.push_back({std::string(nm, strnlen(nm, maxn)), obj});
            }
        }
    }
    std::vector<std::pair<std::string, uint64_t>> children(uint64_t addr) const {
        std::vector<std::pair<std::string, uint64_t>> out;
        for (const Msg &m : messages(addr)) {
            if (m.type == 0x11) {                                         // symbol table: B-tree + local heap
                if (m.n < 2 * O) throw H5Error("corrupt symbol table message");
                const uint64_t bt = off_at(m.p), hp = off_at(m.p + O);
                const uint8_t *h = abs(hp, 8 + 2 * (uint64_t) L + O);
                if (memcmp(h, "HEAP", 4)) throw H5Error("corrupt local heap");
                const uint64_t hsize = len_at(h + 8), hdata = off_at(h + 8 + 2 * L);
                abs(hdata, hsize);
                if (bt != UNDEF) walk_group_btree(bt, hdata, hsize, 0, out);
            } else if (m.type == 0x06) {                                  // link message (compact new-style group)
                if (m.n < 4 || m.p[0] != 1) throw H5Error("unsupported link message");
                const uint8_t fl = m.p[1];
                size_t q = 2;
                unsigned type = 0;
                if (fl & 0x08) type = m.p[q++];
                if (fl & 0x04) q += 8;
                if (fl & 0x10) q += 1;
                const unsigned lsz = 1u << (fl & 3);
                if (q + lsz > m.n) throw H5Error("corrupt link message");
                const uint64_t nlen = le(m.p + q, lsz);
                q += lsz;
                if (q + nlen > m.n) throw H5Error("corrupt link message");
                std::string nm((const char *) m.p + q, (size_t) nlen);
                q += nlen;
                if (type != 0) continue;                                  // soft / external link
                if (q + O > m.n) throw H5Error("corrupt link message");
                out.push_back({nm, off_at(m.p + q)});
            } else if (m.type == 0x02) {                                  // link info
                if (m.n >= 2) {
                    size_t q = 2 + ((m.p[1] & 1) ? 8 : 0);
                    if (q + O <= m.n && off_at(m.p + q) != UNDEF)
                        throw H5Error("unsupported HDF5 feature: densely stored links (fractal heap)");
                }
            }
        }
        return out;
    }
    uint64_t child(uint64_t addr, const std::string &name) const {
        for (auto &c : children(addr)) if (c.first == name) return c.second;
        return UNDEF;
    }
    uint64_t lookup(const std::string &path) const {                     // "/a/b/c"; UNDEF when absent
        uint64_t cur = root;
        size_t i = 0;
        while (i < path.size()) {
            while (i < path.size() && path[i] == '/') i++;
            size_t j = path.find('/', i);
            if (j == std::string::npos) j = path.size();
            if (j > i) {
                cur = child(cur, path.substr(i, j - i));
                if (cur == UNDEF) return UNDEF;
            }
            i = j;
        }
        return cur;
    }

    // ---- datatypes, dataspaces, attributes ----------------------------------------------------------------
    DType datatype(const uint8_t *p, size_t n) const {
        if (n < 8) throw H5Error("corrupt datatype message");
        DType t;
        t.cls = p[0] & 15;
        t.size = (uint32_t) le(p + 4, 4);
        const uint8_t b0 = p[1];
        if (t.cls == 0) { t.big_endian = b0 & 1; t.is_signed = b0 & 8; }
        else if (t.cls == 1) { t.big_endian = b0 & 1; }
        else if (t.cls == 3) { t.str_pad = b0 & 15; }
        else if (t.cls == 9) { t.vlen_string = (b0 & 15) == 1; t.str_pad = (b0 >> 4) & 15; }
        return t;
    }
    DSpace dataspace(const uint8_t *p, size_t n) const {
        if (n < 4) throw H5Error("corrupt dataspace message");
        DSpace s;
        const unsigned ver = p[0];
        s.rank = p[1];
        size_t q;
        if (ver == 1) q = 8;
        else if (ver == 2) { q = 4; s.null_space = p[3] == 2; }
        else throw H5Error("unsupported dataspace version");
        if (s.rank > 4) throw H5Error("unsupported dataspace rank");
        if (q + (size_t) s.rank * L > n) throw H5Error("corrupt dataspace message");
        for (int i = 0; i < s.rank; i++) { s.dims[i] = len_at(p + q + (size_t) i * L); s.count *= s.dims[i]; }
        if (s.null_space) s.count = 0;
        return s;
    }

    // hdf5_tools' String_Reader: the first element of an attribute as text
    std::string value_text(const DType &t, const uint8_t *v, size_t avail) const {
        char buf[64];
        if (t.cls == 3) {
            if (t.size > avail) throw H5Error("corrupt attribute (string overruns the message)");
            std::string s((const char *) v, t.size);
            if (t.str_pad == 0) s.resize(strnlen(s.data(), s.size()));
            else if (t.str_pad == 2) while (!s.empty() && s.back() == ' ') s.pop_back();
            while (!s.empty() && s.back() == '\0') s.pop_back();
            return s;
        }
        if (t.cls == 9 && t.vlen_string) {                                 // length, global heap collection, object index
            if (avail < 8 + (size_t) O) throw H5Error("corrupt attribute (variable-length reference)");
            const uint64_t coll = off_at(v + 4);
            const uint32_t index = (uint32_t) le(v + 4 + O, 4);
            if (coll == UNDEF || index == 0) return std::string();
            const uint8_t *g = abs(coll, 8 + (uint64_t) L);
            if (memcmp(g, "GCOL", 4)) throw H5Error("corrupt global heap collection");
            const uint64_t csize = len_at(g + 8);
            g = abs(coll, csize);
            uint64_t q = 8 + L;
            while (q + 8 + L <= csize) {
                const uint32_t oi = (uint32_t) le(g + q, 2);
                const uint64_t osz = len_at(g + q + 8);
                if (oi == 0) break;
                if (osz > csize - (q + 8 + L)) throw H5Error("corrupt global heap object");     // subtraction form: no wrap-around
                if (oi == index) {
                    const char *s = (const char *) g + q + 8 + L;
                    return std::string(s, strnlen(s, (size_t) osz));
                }
                q += 8 + L + ((osz + 7) & ~7ull);
            }
            throw H5Error("global heap object not found");
        }
        if (t.cls == 0) {
            if (t.size > 8 || t.size == 0 || t.size > avail) throw H5Error("unsupported integer attribute size");
            uint64_t u = 0;
            for (uint32_t i = 0; i < t.size; i++) u |= (uint64_t) v[t.big_endian ? t.size - 1 - i : i] << (8 * i);
            if (t.is_signed) {
                long long s = (long long) u;
                if (t.size < 8 && (u >> (8 * t.size - 1)) & 1) s = (long long) (u | (~0ull << (8 * t.size)));
                snprintf(buf, sizeof buf, "%lld", s);
            } else {
                snprintf(buf, sizeof buf, "%llu", (unsigned long long) u);
            }
            return buf;
        }
        if (t.cls == 1) {
            if ((t.size != 4 && t.size != 8) || t.size > avail) throw H5Error("unsupported floating-point attribute size");
            uint8_t raw[8];
            for (uint32_t i = 0; i < t.size; i++) raw[i] = v[t.big_endian ? t.size - 1 - i : i];
            double x;
            if (t.size == 4) { float f; memcpy(&f, raw, 4); x = f; } else memcpy(&x, raw, 8);
            snprintf(buf, sizeof buf, "%g", x);                          // ostringstream << double: 6 significant digits
            return buf;
        }
        throw H5Error("unsupported attribute datatype class");
    }

    // File::get_attr_map: every attribute of the object as (name, text)
    std::vector<std::pair<std::string, std::string>> attributes(uint64_t addr) const {
        std::vector<std::pair<std::string, std::string>> out;
        for (const Msg &m : messages(addr)) {
            if (m.type == 0x15) {                                         // attribute info: dense storage?
                if (m.n >= 2) {
                    size_t q = 2 + ((m.p[1] & 1) ? 2 : 0);
                    if (q + O <= m.n && off_at(m.p + q) != UNDEF)
                        throw H5Error("unsupported HDF5 feature: densely stored attributes (fractal heap)");
                }
                continue;
            }
            if (m.type != 0x0C) continue;
            if (m.flags & 2) throw H5Error("unsupported HDF5 feature: shared attribute message");
            if (m.n < 8) throw H5Error("corrupt attribute message");
            const unsigned ver = m.p[0];
            const size_t nsz = (size_t) le(m.p + 2, 2), tsz = (size_t) le(m.p + 4, 2), ssz = (size_t) le(m.p + 6, 2);
            size_t q = 8;
            if (ver == 3) q = 9;
            else if (ver != 1 && ver != 2) throw H5Error("unsupported attribute message version");
            if (ver != 1 && (m.p[1] & 3)) throw H5Error("unsupported HDF5 feature: shared attribute datatype/dataspace");
            auto pad = [&](size_t x) { return ver == 1 ? (x + 7) & ~(size_t) 7 : x; };
            if (q + pad(nsz) + pad(tsz) + pad(ssz) > m.n) throw H5Error("corrupt attribute message");
            std::string name((const char *) m.p + q, strnlen((const char *) m.p + q, nsz));
            q += pad(nsz);
            DType t = datatype(m.p + q, tsz);
            q += pad(tsz);
            DSpace s = dataspace(m.p + q, ssz);
            q += pad(ssz);
            if (s.count == 0) { out.push_back({name, std::string()}); continue; }
            out.push_back({name, value_text(t, m.p + q, m.n - q)});
        }
        return out;
    }

    // ---- datasets ------------------------------------------------------------------------------------------
    Dataset dataset(uint64_t addr) const {
        Dataset D;
        bool have_t = false, have_s = false;
        for (const Msg &m : messages(addr)) {
            if (m.type == 0x03) {
                if (m.flags & 2) throw H5Error("unsupported HDF5 feature: committed (shared) datatype");
                D.dt = datatype(m.p, m.n); have_t = true;
            } else if (m.type == 0x01) {
                D.ds = dataspace(m.p, m.n); have_s = true;
            } else if (m.type == 0x0B) {
                if (m.n < 2) throw H5Error("corrupt filter pipeline message");
                const unsigned ver = m.p[0], nf = m.p[1];
                size_t q = ver == 1 ? 8 : 2;
                if (ver != 1 && ver != 2) throw H5Error("unsupported filter pipeline version");
                for (unsigned i = 0; i < nf; i++) {
                    if (q + 6 > m.n) throw H5Error("corrupt filter pipeline message");
                    Filter f;
                    f.id = (uint16_t) le(m.p + q, 2);
                    size_t name_len = 0;
                    if (ver == 1 || f.id >= 256) { name_len = (size_t) le(m.p + q + 2, 2); q += 4; } else q += 2;
                    if (q + 4 > m.n) throw H5Error("corrupt filter pipeline message");
                    const unsigned ncd = (unsigned) le(m.p + q + 2, 2);
                    q += 4;
                    q += ver == 1 ? ((name_len + 7) & ~(size_t) 7) : name_len;
                    if (q + 4 * (size_t) ncd > m.n) throw H5Error("corrupt filter pipeline message");
                    for (unsigned k = 0; k < ncd; k++) f.cd.push_back((uint32_t) le(m.p + q + 4 * k, 4));
                    q += 4 * (size_t) ncd;
                    if (ver == 1 && (ncd & 1)) q += 4;
                    D.filters.push_back(f);
                }
            } else if (m.type == 0x08) {
                if (m.n < 2) throw H5Error("corrupt data layout message");
                const unsigned ver = m.p[0];
                if (ver == 3) {
                    D.layout = m.p[1];
                    if (D.layout == 0) {
                        D.compact_size = (size_t) le(m.p + 2, 2);
                        if (4 + D.compact_size > m.n) throw H5Error("corrupt compact layout");
                        D.compact = m.p + 4;
                    } else if (D.layout == 1) {
                        if (2 + (size_t) O + L > m.n) throw H5Error("corrupt contiguous layout");
                        D.addr = off_at(m.p + 2); D.size = len_at(m.p + 2 + O);
                    } else if (D.layout == 2) {
                        D.chunk_rank = m.p[2];
                        if (3 + (size_t) O + 4 * (size_t) D.chunk_rank > m.n || D.chunk_rank < 2) throw H5Error("corrupt chunked layout");
                        D.addr = off_at(m.p + 3);
                        D.chunk_elems = le(m.p + 3 + O, 4);
                    } else throw H5Error("unsupported data layout class");
                } else if (ver == 1 || ver == 2) {
                    const unsigned dim = m.p[1];
                    D.layout = m.p[2];
                    size_t q = 8;
                    if (D.layout != 0) { if (q + O > m.n) throw H5Error("corrupt data layout message"); D.addr = off_at(m.p + q); q += O; }
                    if (q + 4 * (size_t) dim > m.n) throw H5Error("corrupt data layout message");
                    if (D.layout == 2) { D.chunk_rank = (int) dim; D.chunk_elems = le(m.p + q, 4); }
                    q += 4 * (size_t) dim;
                    if (D.layout == 0) {
                        if (q + 4 > m.n) throw H5Error("corrupt compact layout");
                        D.compact_size = (size_t) le(m.p + q, 4);
                        if (q + 4 + D.compact_size > m.n) throw H5Error("corrupt compact layout");
                        D.compact = m.p + q + 4;
                    }
                } else {
                    throw H5Error("unsupported HDF5 feature: data layout message version 4 (libver 'latest' chunk indexes)");
                }
            }
        }
        if (!have_t || !have_s || D.layout < 0) throw H5Error("not a dataset");
        return D;
    }

    static void unfilter(const Dataset &D, uint32_t mask, std::vector<uint8_t> &buf) {
        for (int i = (int) D.filters.size() - 1; i >= 0; i--) {
            if ((mask >> i) & 1u) continue;
            const Filter &f = D.filters[i];
            if (f.id == 1) {                                              // deflate
                // a chunk inflates to chunk_elems elements (+ the fletcher32 word when that filter ran first)
                std::vector<uint8_t> out((size_t) D.chunk_elems * D.dt.size + 64);
                uLongf n = (uLongf) out.size();
                if (uncompress(out.data(), &n, buf.data(), (uLong) buf.size()) != Z_OK) throw H5Error("deflate: corrupt chunk");
                out.resize(n);
                buf.swap(out);
            } else if (f.id == 2) {                                       // shuffle
                const size_t es = f.cd.empty() ? D.dt.size : f.cd[0];
                if (es > 1 && buf.size() >= es) {
                    const size_t n = buf.size() / es;
                    std::vector<uint8_t> out(buf.size());
                    for (size_t b = 0; b < es; b++)
                        for (size_t k = 0; k < n; k++) out[k * es + b] = buf[b * n + k];
                    memcpy(out.data() + n * es, buf.data() + n * es, buf.size() - n * es);
                    buf.swap(out);
                }
            } else if (f.id == 3) {                                       // fletcher32: checksum appended
                if (buf.size() < 4) throw H5Error("fletcher32: corrupt chunk");
                buf.resize(buf.size() - 4);
            } else if (f.id == 32020) {
                throw H5Error("unsupported HDF5 filter 32020 (VBZ): convert the file to gzip compression "
                              "(e.g. ont_fast5_api's compress_fast5 -c gzip)");
            } else {
                throw H5Error("unsupported HDF5 filter id " + std::to_string(f.id));
            }
        }
    }

    void walk_chunks(const Dataset &D, uint64_t node, int depth, int16_t *dst, uint64_t want) const {
        if (depth > 32) throw H5Error("corrupt chunk B-tree");
        const uint8_t *p = abs(node, 8 + 2 * (uint64_t) O);
        if (memcmp(p, "TREE", 4) || p[4] != 1) throw H5Error("corrupt chunk B-tree node");
        const unsigned level = p[5], used = (unsigned) le(p + 6, 2);
        const uint64_t ksz = 8 + 8 * (uint64_t) D.chunk_rank;
        const uint8_t *e = abs(node + 8 + 2 * O, used * (ksz + O) + ksz);
        for (unsigned i = 0; i < used; i++) {
            const uint8_t *key = e + i * (ksz + O);
            const uint64_t child = off_at(key + ksz);
            const uint64_t first = le(key + 8, 8);                        // offset of the chunk along dimension 0
            if (first >= want) continue;                                  // beyond the requested prefix
            if (level > 0) { walk_chunks(D, child, depth + 1, dst, want); continue; }
            const uint32_t nbytes = (uint32_t) le(key, 4), mask = (uint32_t) le(key + 4, 4);
            const uint8_t *raw = abs(child, nbytes);
            std::vector<uint8_t> buf(raw, raw + nbytes);
            unfilter(D, mask, buf);
            const uint64_t n = std::min<uint64_t>(std::min<uint64_t>(D.chunk_elems, buf.size() / 2), want - first);
            memcpy(dst + first, buf.data(), n * 2);
        }
    }

    // the first `want` elements of a one-dimensional little-endian int16 dataset
    void read_i16(const Dataset &D, int16_t *dst, uint64_t want) const {
        if (D.dt.cls != 0 || D.dt.size != 2 || D.dt.big_endian) throw H5Error("Signal is not a little-endian 16-bit integer dataset");
        if (D.ds.rank != 1) throw H5Error("Signal is not one-dimensional");
        want = std::min(want, D.ds.count);
        if (!want) return;
        memset(dst, 0, want * 2);                                         // unallocated storage reads as the fill value 0
        if (D.layout == 0) {
            memcpy(dst, D.compact, std::min<uint64_t>(want * 2, D.compact_size));
        } else if (D.layout == 1) {
            if (D.addr == UNDEF) return;
            if (!D.filters.empty()) throw H5Error("filtered contiguous dataset");
            memcpy(dst, abs(D.addr, want * 2), want * 2);
        } else {
            if (D.chunk_rank != 2 || !D.chunk_elems) throw H5Error("unsupported chunk shape");
            if (D.addr != UNDEF) walk_chunks(D, D.addr, 0, dst, want);
        }
    }
};

struct ReadEntry {
    std::string raw_path, ch_path;
    uint64_t raw_addr, ch_addr, sig_addr;
};

}  // namespace

struct unc_fast5 {
    H5File f;
    bool single = false;
    std::vector<ReadEntry> reads;
    std::vector<std::string> ids;       // storage behind unc_fast5_read::read_id
};

static thread_local std::string g_f5_err;

// atoi as the reference applies it to attribute text (src/read_buffer.cpp:204-214): strtol narrowed to int
static int32_t text_atoi(const std::string &s) { return (int32_t) strtol(s.c_str(), nullptr, 10); }

static void fill_info(unc_fast5 *h, uint32_t i, unc_fast5_read *r) {
    const ReadEntry &e = h->reads[i];
    memset(r, 0, sizeof *r);
    float cal_digit = 1, cal_range = 1, cal_offset = 0;                   // defaults of src/read_buffer.cpp:212
    for (auto &a : h->f.attributes(e.raw_addr)) {
        if (a.first == "read_id") h->ids[i] = a.second;
        else if (a.first == "read_number") r->number = text_atoi(a.second);
        else if (a.first == "start_time") r->start_sample = text_atoi(a.second);
    }
    for (auto &a : h->f.attributes(e.ch_addr)) {
        if (a.first == "channel_number") r->channel = text_atoi(a.second);
        else if (a.first == "digitisation") cal_digit = (float) atof(a.second.c_str());
        else if (a.first == "range") cal_range = (float) atof(a.second.c_str());
        else if (a.first == "offset") cal_offset = (float) atof(a.second.c_str());
    }
    r->read_id = h->ids[i].c_str();
    r->cal_digitisation = cal_digit; r->cal_range = cal_range; r->cal_offset = cal_offset;
    Dataset D = h->f.dataset(e.sig_addr);
    if (D.ds.rank != 1) throw H5Error("Signal is not one-dimensional");
    r->n_samples = D.ds.count;
}

extern "C" {

const char *unc_fast5_last_error(void) { return g_f5_err.c_str(); }

int unc_fast5_open(const char *path, unc_fast5 **out) {
    if (!path || !out) { g_f5_err = "null argument"; return UNC_E_ARG; }
    unc_fast5 *h = new unc_fast5();
    try {
        h->f.open(path);
        auto top = h->f.children(h->f.root);
        for (auto &c : top) if (c.first == "Raw") h->single = true;       // Fast5Reader::open_next (src/fast5_reader.cpp:134-150)
        if (h->single) {
            const uint64_t reads = h->f.lookup("/Raw/Reads"), ch = h->f.lookup("/UniqueGlobalKey/channel_id");
            if (reads == UNDEF || ch == UNDEF) throw H5Error("single-read fast5 without /Raw/Reads or /UniqueGlobalKey/channel_id");
            for (auto &c : h->f.children(reads)) {
                ReadEntry e{"/Raw/Reads/" + c.first, "/UniqueGlobalKey/channel_id", c.second, ch, h->f.child(c.second, "Signal")};
                if (e.sig_addr == UNDEF) throw H5Error(e.raw_path + ": no Signal dataset");
                h->reads.push_back(e);
            }
        } else {
            for (auto &c : top) {
                const uint64_t raw = h->f.child(c.second, "Raw"), ch = h->f.child(c.second, "channel_id");
                if (raw == UNDEF || ch == UNDEF) continue;                // not a read group
                ReadEntry e{"/" + c.first + "/Raw", "/" + c.first + "/channel_id", raw, ch, h->f.child(raw, "Signal")};
                if (e.sig_addr == UNDEF) throw H5Error(e.raw_path + ": no Signal dataset");
                h->reads.push_back(e);
            }
        }
        h->ids.resize(h->reads.size());
    } catch (const std::exception &e) {
        g_f5_err = std::string(path) + ": " + e.what();
        delete h;
        return UNC_E_IO;
    }
    *out = h;
    return UNC_OK;
}

int unc_fast5_count(const unc_fast5 *h, uint32_t *n_reads, int *single_read_format) {
    if (!h || !n_reads) { g_f5_err = "null argument"; return UNC_E_ARG; }
    *n_reads = (uint32_t) h->reads.size();
    if (single_read_format) *single_read_format = h->single ? 1 : 0;
    return UNC_OK;
}

int unc_fast5_info(unc_fast5 *h, uint32_t i, unc_fast5_read *info) {
    if (!h || !info || i >= h->reads.size()) { g_f5_err = "bad argument"; return UNC_E_ARG; }
    try { fill_info(h, i, info); }
    catch (const std::exception &e) { g_f5_err = h->reads[i].raw_path + ": " + e.what(); return UNC_E_IO; }
    return UNC_OK;
}

int unc_fast5_load(unc_fast5 *h, uint32_t first, uint32_t n, uint64_t max_samples_per_read, int16_t *dst,
                   uint64_t capacity, unc_fast5_read *info, int threads) {
    if (!h || !info || (!dst && capacity) || first > h->reads.size() || n > h->reads.size() - first) {
        g_f5_err = "bad argument";
        return UNC_E_ARG;
    }
    try {
        unsigned nt = threads > 0 ? (unsigned) threads : std::max(1u, std::thread::hardware_concurrency());
        nt = std::min<unsigned>(nt, std::max<uint32_t>(n, 1u));
        std::vector<std::string> errs(nt);
        // two parallel phases over the reads, handed out by an atomic counter: (1) attributes + signal length,
        // (2) after the prefix sum of the lengths, the signals themselves
        auto run = [&](const std::function<void(uint32_t)> &fn) {
            std::atomic<uint32_t> next(0);
            auto work = [&](unsigned t) {
                uint32_t i = 0;
                try { while ((i = next.fetch_add(1)) < n) fn(i); }
                catch (const std::exception &e) { errs[t] = h->reads[first + i].raw_path + ": " + e.what(); }
            };
            std::vector<std::thread> pool;
            for (unsigned t = 1; t < nt; t++) pool.emplace_back(work, t);
            work(0);
            for (auto &t : pool) t.join();
            for (auto &e : errs) if (!e.empty()) throw H5Error(e);
        };
        run([&](uint32_t i) {
            fill_info(h, first + i, &info[i]);
            if (max_samples_per_read && info[i].n_samples > max_samples_per_read) info[i].n_samples = max_samples_per_read;
        });
        uint64_t total = 0;
        for (uint32_t i = 0; i < n; i++) { info[i].sample_offset = total; total += info[i].n_samples; }
        if (total > capacity) { g_f5_err = "staging buffer too small"; return UNC_E_TOO_LARGE; }
        run([&](uint32_t i) {
            Dataset D = h->f.dataset(h->reads[first + i].sig_addr);
            h->f.read_i16(D, dst + info[i].sample_offset, info[i].n_samples);
        });
    } catch (const std::exception &e) {
        g_f5_err = e.what();
        return UNC_E_IO;
    }
    return UNC_OK;
}

void unc_fast5_close(unc_fast5 *h) { delete h; }

}  // extern "C"
