/* Writes small fast5 files with the REAL libhdf5 (the reference's vendored 1.8.21, built under /tmp by build.sh)
 * so that the product's own HDF5 subset reader (uncalled_b200/csrc/unc_fast5.cpp) is tested against files it did
 * not write.  Each file exercises different parts of the file format; see the table in main().
 * Test tooling only -- never linked into the product. */
#include <hdf5.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t) (rng_state >> 11); }

enum { STR_VLEN, STR_NULLTERM, STR_NULLPAD, STR_SPACEPAD };
enum { LAY_CHUNK_GZIP, LAY_CHUNK_SHUF_GZIP_FLETCHER, LAY_CHUNK_PLAIN, LAY_CONTIG, LAY_COMPACT };

static void attr_str(hid_t obj, const char *name, const char *val, int kind) {
    hid_t t = H5Tcopy(H5T_C_S1), s = H5Screate(H5S_SCALAR), a;
    if (kind == STR_VLEN) {
        H5Tset_size(t, H5T_VARIABLE);
        H5Tset_cset(t, H5T_CSET_UTF8);
        a = H5Acreate2(obj, name, t, s, H5P_DEFAULT, H5P_DEFAULT);
        H5Awrite(a, t, &val);
    } else {
        size_t n = strlen(val) + (kind == STR_NULLTERM ? 1 : 5);
        char *buf = (char *) calloc(n + 1, 1);
        H5Tset_size(t, n);
        H5Tset_strpad(t, kind == STR_NULLTERM ? H5T_STR_NULLTERM : kind == STR_NULLPAD ? H5T_STR_NULLPAD : H5T_STR_SPACEPAD);
        memset(buf, kind == STR_SPACEPAD ? ' ' : 0, n);
        memcpy(buf, val, strlen(val));
        a = H5Acreate2(obj, name, t, s, H5P_DEFAULT, H5P_DEFAULT);
        H5Awrite(a, t, buf);
        free(buf);
    }
    H5Aclose(a); H5Sclose(s); H5Tclose(t);
}
static void attr_num(hid_t obj, const char *name, hid_t file_type, hid_t mem_type, const void *val) {
    hid_t s = H5Screate(H5S_SCALAR), a = H5Acreate2(obj, name, file_type, s, H5P_DEFAULT, H5P_DEFAULT);
    H5Awrite(a, mem_type, val);
    H5Aclose(a); H5Sclose(s);
}

typedef struct { int n_reads, min_len, max_len, str_kind, layout, chunk, single, latest, big_start, float32_cal, n_extra_attrs; } Spec;

static void signal_data(int16_t *x, int n, int salt) {
    int level = 400 + (int) (rnd() % 200);
    for (int i = 0; i < n; i++) {
        if (rnd() % 9 == 0) level = 350 + (int) (rnd() % 400);
        x[i] = (int16_t) (level + (int) (rnd() % 21) - 10);
    }
    if (n > 10) { x[3] = -5; x[n / 2] = (int16_t) -32768; x[n - 2] = 32767; x[7] = (int16_t) (salt & 1 ? -1 : 0); }   /* negative DAC values: the u16 reinterpretation */
}

static void write_read(hid_t parent_raw, hid_t ch, const Spec *sp, int idx, const char *uuid) {
    int n = sp->min_len + (int) (rnd() % (unsigned) (sp->max_len - sp->min_len + 1));
    if (idx == 1) n = sp->min_len;                      /* multi_contig: an empty Signal */
    int16_t *x = (int16_t *) malloc(sizeof(int16_t) * (size_t) (n ? n : 1));
    signal_data(x, n, idx);
    hsize_t dims[1] = {(hsize_t) n}, maxd[1] = {H5S_UNLIMITED}, cdims[1] = {(hsize_t) sp->chunk};
    hid_t dcpl = H5Pcreate(H5P_DATASET_CREATE);
    hid_t space;
    if (sp->layout <= LAY_CHUNK_PLAIN) {
        space = H5Screate_simple(1, dims, maxd);
        H5Pset_chunk(dcpl, 1, cdims);
        if (sp->layout == LAY_CHUNK_SHUF_GZIP_FLETCHER) { H5Pset_shuffle(dcpl); H5Pset_deflate(dcpl, 4); H5Pset_fletcher32(dcpl); }
        else if (sp->layout == LAY_CHUNK_GZIP) H5Pset_deflate(dcpl, 1);
    } else {
        space = H5Screate_simple(1, dims, NULL);
        H5Pset_layout(dcpl, sp->layout == LAY_CONTIG ? H5D_CONTIGUOUS : H5D_COMPACT);
    }
    hid_t d = H5Dcreate2(parent_raw, "Signal", H5T_STD_I16LE, space, H5P_DEFAULT, dcpl, H5P_DEFAULT);
    if (n) H5Dwrite(d, H5T_NATIVE_INT16, H5S_ALL, H5S_ALL, H5P_DEFAULT, x);
    H5Dclose(d); H5Sclose(space); H5Pclose(dcpl);
    free(x);

    int32_t number = 100 + idx * 7;
    uint64_t start = sp->big_start ? 4294967296ull * (uint64_t) (idx % 3) + 3000000000ull + (uint64_t) idx * 4001u : 1000u + (uint64_t) idx * 4001u;
    uint32_t duration = (uint32_t) n;
    attr_str(parent_raw, "read_id", uuid, sp->str_kind);
    attr_num(parent_raw, "read_number", H5T_STD_I32LE, H5T_NATIVE_INT32, &number);
    attr_num(parent_raw, "start_time", H5T_STD_U64LE, H5T_NATIVE_UINT64, &start);
    attr_num(parent_raw, "duration", H5T_STD_U32LE, H5T_NATIVE_UINT32, &duration);
    for (int k = 0; k < sp->n_extra_attrs; k++) {      /* pushes the header into continuation blocks */
        char nm[32];
        double v = k * 1.5;
        snprintf(nm, sizeof nm, "extra_attribute_%02d", k);
        attr_num(parent_raw, nm, H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &v);
    }
    char chs[16];
    snprintf(chs, sizeof chs, "%d", 1 + (idx * 37) % 512);
    double digit = 8192.0, range = 1467.61 + 0.123456789 * idx, offset = (double) (5 + idx % 11) - (idx % 2 ? 0.0 : 11.0), rate = 4000.0;
    attr_str(ch, "channel_number", chs, sp->str_kind);
    if (sp->float32_cal) {
        float fd = (float) digit, fr = (float) range, fo = (float) offset;
        attr_num(ch, "digitisation", H5T_IEEE_F32LE, H5T_NATIVE_FLOAT, &fd);
        attr_num(ch, "range", H5T_IEEE_F32BE, H5T_NATIVE_FLOAT, &fr);
        attr_num(ch, "offset", H5T_STD_I16BE, H5T_NATIVE_FLOAT, &fo);
    } else {
        attr_num(ch, "digitisation", H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &digit);
        attr_num(ch, "range", H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &range);
        attr_num(ch, "offset", H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &offset);
    }
    attr_num(ch, "sampling_rate", H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &rate);
}

static void make(const char *path, const Spec *sp) {
    hid_t fapl = H5Pcreate(H5P_FILE_ACCESS);
    if (sp->latest) H5Pset_libver_bounds(fapl, H5F_LIBVER_LATEST, H5F_LIBVER_LATEST);
    hid_t fcpl = H5Pcreate(H5P_FILE_CREATE);
    if (sp->n_reads > 50) H5Pset_sym_k(fcpl, 2, 2);     /* tiny group B-tree nodes: a three-level tree with few reads */
    hid_t f = H5Fcreate(path, H5F_ACC_TRUNC, fcpl, fapl);
    H5Pclose(fcpl);
    attr_str(f, "file_version", sp->single ? "1.0" : "2.0", sp->str_kind);
    if (sp->single) {
        hid_t raw = H5Gcreate2(f, "Raw", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        hid_t reads = H5Gcreate2(raw, "Reads", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        hid_t rd = H5Gcreate2(reads, "Read_101", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        hid_t ugk = H5Gcreate2(f, "UniqueGlobalKey", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        hid_t ch = H5Gcreate2(ugk, "channel_id", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        hid_t trk = H5Gcreate2(ugk, "tracking_id", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        attr_str(trk, "run_id", "0123456789abcdef", sp->str_kind);
        write_read(rd, ch, sp, 0, "0a1b2c3d-0000-4000-8000-00000000beef");
        H5Gclose(trk); H5Gclose(ch); H5Gclose(ugk); H5Gclose(rd); H5Gclose(reads); H5Gclose(raw);
    } else {
        for (int i = 0; i < sp->n_reads; i++) {
            char uuid[64], name[80];
            snprintf(uuid, sizeof uuid, "%08x-%04x-4%03x-a%03x-%012llx", rnd(), rnd() & 0xffff, rnd() & 0xfff, rnd() & 0xfff,
                     (unsigned long long) (((uint64_t) rnd() << 20) ^ rnd()) & 0xffffffffffffull);
            snprintf(name, sizeof name, "read_%s", uuid);
            hid_t g = H5Gcreate2(f, name, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
            attr_str(g, "run_id", "0123456789abcdef", sp->str_kind);
            hid_t raw = H5Gcreate2(g, "Raw", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
            hid_t ch = H5Gcreate2(g, "channel_id", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
            hid_t ctx = H5Gcreate2(g, "context_tags", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
            write_read(raw, ch, sp, i, uuid);
            H5Gclose(ctx); H5Gclose(ch); H5Gclose(raw); H5Gclose(g);
        }
    }
    H5Fclose(f); H5Pclose(fapl);
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s OUT_DIR\n", argv[0]); return 1; }
    static const struct { const char *name; Spec sp; } files[] = {
        /* name                      n  min    max     strings       layout                        chunk single latest big f32 extra */
        {"multi_gzip.fast5",        {12, 3000, 9000,   STR_VLEN,     LAY_CHUNK_GZIP,               1000, 0, 0, 0, 0, 0}},   /* ont_fast5_api-like: vlen UTF-8 strings (global heap) */
        {"multi_deep_chunks.fast5", {3, 70000, 90000,  STR_NULLTERM, LAY_CHUNK_SHUF_GZIP_FLETCHER, 500,  0, 0, 1, 0, 0}},   /* > 64 chunks: two-level chunk B-tree; shuffle + fletcher32; start_time > 2^32 */
        {"multi_many_reads.fast5",  {70, 40,   90,     STR_NULLPAD,  LAY_CHUNK_PLAIN,              64,   0, 0, 0, 0, 0}},   /* 70 root links with group K = 2: three-level group B-tree; unfiltered chunks */
        {"multi_contig.fast5",      {5, 0,     5000,   STR_SPACEPAD, LAY_CONTIG,                   1,    0, 0, 0, 1, 12}},  /* contiguous, f32 / big-endian / integer calibration attrs, header continuation blocks */
        {"multi_compact.fast5",     {4, 100,   3000,   STR_VLEN,     LAY_COMPACT,                  1,    0, 0, 0, 0, 0}},   /* compact layout (data inside the object header) */
        {"multi_latest.fast5",      {6, 2000,  6000,   STR_VLEN,     LAY_CHUNK_GZIP,               700,  0, 1, 0, 0, 0}},   /* libver latest: superblock v2, v2 object headers, link messages */
        {"single_gzip.fast5",       {1, 20000, 20000,  STR_NULLTERM, LAY_CHUNK_GZIP,               4096, 1, 0, 0, 0, 0}},   /* MinKNOW single-read layout */
    };
    for (size_t i = 0; i < sizeof files / sizeof files[0]; i++) {
        char path[4096];
        snprintf(path, sizeof path, "%s/%s", argv[1], files[i].name);
        rng_state = 88172645463325252ull + i * 7919u;
        make(path, &files[i].sp);
        printf("wrote %s\n", path);
    }
    return 0;
}
