/* Writes ONE multi-read fast5 for the fast5 -> GPU throughput measurement (configs[2]/[3] of BASELINE.json, SURVEY.md 8(d):
 * `/read_<id>/Raw{Signal; attrs read_id, read_number, start_time}`, `/read_<id>/channel_id{channel_number, digitisation,
 * offset, range}`, gzip-compressed int16 signal, 4000 reads per file) with the REAL libhdf5 (the reference's vendored 1.8.21,
 * built under /tmp by build.sh), from int16 DAC samples made by tools/make_bench_fast5.py.
 *   make_bench_fast5 OUT.fast5 RAW_I16_FILE N_READS N_SAMPLES
 * Test / bench tooling only -- never linked into the product (which READS fast5 with its own HDF5 subset reader). */
#include <hdf5.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void attr_str(hid_t obj, const char *name, const char *val) {
    hid_t t = H5Tcopy(H5T_C_S1), s = H5Screate(H5S_SCALAR);
    H5Tset_size(t, H5T_VARIABLE);
    H5Tset_cset(t, H5T_CSET_UTF8);
    hid_t a = H5Acreate2(obj, name, t, s, H5P_DEFAULT, H5P_DEFAULT);
    H5Awrite(a, t, &val);
    H5Aclose(a); H5Sclose(s); H5Tclose(t);
}
static void attr_num(hid_t obj, const char *name, hid_t file_type, hid_t mem_type, const void *val) {
    hid_t s = H5Screate(H5S_SCALAR), a = H5Acreate2(obj, name, file_type, s, H5P_DEFAULT, H5P_DEFAULT);
    H5Awrite(a, mem_type, val);
    H5Aclose(a); H5Sclose(s);
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s OUT.fast5 RAW_I16 N_READS N_SAMPLES\n", argv[0]); return 1; }
    const int n_reads = atoi(argv[3]), n_samples = atoi(argv[4]);
    FILE *in = fopen(argv[2], "rb");
    if (!in) { perror(argv[2]); return 1; }
    int16_t *x = (int16_t *) malloc(sizeof(int16_t) * (size_t) n_samples);
    hid_t f = H5Fcreate(argv[1], H5F_ACC_TRUNC, H5P_DEFAULT, H5P_DEFAULT);
    attr_str(f, "file_version", "2.0");
    for (int i = 0; i < n_reads; i++) {
        if (fread(x, 2, (size_t) n_samples, in) != (size_t) n_samples) { fprintf(stderr, "short raw file\n"); return 1; }
        char uuid[64], name[80], chs[16];
        snprintf(uuid, sizeof uuid, "%08x-b200-4000-a000-%012x", 0x5eed0000u + (unsigned) i, (unsigned) i);
        snprintf(name, sizeof name, "read_%s", uuid);
        hid_t g = H5Gcreate2(f, name, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        attr_str(g, "run_id", "b200benchb200bench");
        hid_t raw = H5Gcreate2(g, "Raw", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        hid_t ch = H5Gcreate2(g, "channel_id", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        hsize_t dims[1] = {(hsize_t) n_samples}, maxd[1] = {H5S_UNLIMITED}, cdims[1] = {(hsize_t) n_samples};
        hid_t dcpl = H5Pcreate(H5P_DATASET_CREATE), space = H5Screate_simple(1, dims, maxd);
        H5Pset_chunk(dcpl, 1, cdims);
        H5Pset_deflate(dcpl, 1);
        hid_t d = H5Dcreate2(raw, "Signal", H5T_STD_I16LE, space, H5P_DEFAULT, dcpl, H5P_DEFAULT);
        H5Dwrite(d, H5T_NATIVE_INT16, H5S_ALL, H5S_ALL, H5P_DEFAULT, x);
        H5Dclose(d); H5Sclose(space); H5Pclose(dcpl);
        int32_t number = i + 1;
        uint64_t start = 1000u + (uint64_t) i * 4001u;
        uint32_t duration = (uint32_t) n_samples;
        attr_str(raw, "read_id", uuid);
        attr_num(raw, "read_number", H5T_STD_I32LE, H5T_NATIVE_INT32, &number);
        attr_num(raw, "start_time", H5T_STD_U64LE, H5T_NATIVE_UINT64, &start);
        attr_num(raw, "duration", H5T_STD_U32LE, H5T_NATIVE_UINT32, &duration);
        snprintf(chs, sizeof chs, "%d", 1 + i % 512);
        double digit = 8192.0, range = 1467.61, offset = 10.0, rate = 4000.0;   /* SURVEY.md 8(d) */
        attr_str(ch, "channel_number", chs);
        attr_num(ch, "digitisation", H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &digit);
        attr_num(ch, "range", H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &range);
        attr_num(ch, "offset", H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &offset);
        attr_num(ch, "sampling_rate", H5T_IEEE_F64LE, H5T_NATIVE_DOUBLE, &rate);
        H5Gclose(ch); H5Gclose(raw); H5Gclose(g);
    }
    H5Fclose(f);
    fclose(in);
    free(x);
    return 0;
}
