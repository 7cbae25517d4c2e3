// Dumps what the REFERENCE's own Fast5Reader / ReadBuffer(hdf5_tools::File&, raw_path, ch_path)
// (src/fast5_reader.cpp, src/read_buffer.cpp, compiled from /root/reference against its vendored libhdf5)
// deliver for a fast5 file: one JSON object per read.  Test tooling for tests/golden/fast5/golden.json.
#include <zlib.h>
#include <cstdio>
#include <string>
#include "fast5_reader.hpp"

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s MAX_CHUNKS FAST5...\n", argv[0]); return 1; }
    ReadBuffer::PRMS.max_chunks = (u32) atoi(argv[1]);
    printf("[\n");
    bool first = true;
    for (int a = 2; a < argc; a++) {
        Fast5Reader rd;
        rd.add_fast5(argv[a]);
        std::string base = argv[a];
        base = base.substr(base.find_last_of('/') + 1);
        while (!rd.empty()) {
            if (rd.buffer_size() == 0 && rd.fill_buffer() == 0) break;
            ReadBuffer r = rd.pop_read();
            const std::vector<float> &s = r.get_raw();
            unsigned long crc = crc32(0L, Z_NULL, 0);
            crc = crc32(crc, (const Bytef *) s.data(), (uInt) (s.size() * sizeof(float)));
            printf("%s {\"file\": \"%s\", \"max_chunks\": %u, \"id\": \"%s\", \"number\": %u, \"start\": %llu, \"channel\": %u, "
                   "\"n\": %u, \"crc32_f32\": %lu, \"head\": [", first ? "" : ",\n", base.c_str(), ReadBuffer::PRMS.max_chunks,
                   r.get_id().c_str(), r.get_number(), (unsigned long long) r.get_start(), (unsigned) r.get_channel(), r.size(), crc);
            for (size_t i = 0; i < s.size() && i < 4; i++) printf("%s%.9g", i ? ", " : "", s[i]);
            printf("]}");
            first = false;
        }
    }
    printf("\n]\n");
    return 0;
}
