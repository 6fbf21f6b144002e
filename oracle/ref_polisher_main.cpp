/*
 * ref_polisher_main.cpp — TEST INFRASTRUCTURE ONLY: command-line front end of ref_polisher_harness.cpp
 * (a separate process, because the reference's parsers use the system zlib, which clashes with the zlib
 * statically linked into the Python interpreter when loaded through ctypes).
 * usage: refpol_dump <reads> <overlaps> <target> <fragment 0|1> <w> <q> <e> <trim> <m> <x> <g> <threads> <out.bin>
 * Output: little-endian, u64 counts [windows, sequences, bases, polished], then the flat arrays in the order
 * written below.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

extern "C" {
void* ref_polisher_open(const char*, const char*, const char*, int, uint32_t, double, double, int, int8_t, int8_t,
                        int8_t, uint32_t);
void ref_polisher_counts(void*, uint64_t*);
void ref_polisher_export(void*, char*, char*, uint64_t*, uint8_t*, uint32_t*, uint32_t*, uint32_t*, uint8_t*, uint64_t*,
                         uint32_t*);
uint32_t ref_polisher_polish(void*);
uint32_t ref_polisher_window_consensus(void*, uint32_t, char*, uint32_t);
uint64_t ref_polisher_polished(void*, uint32_t, char*, uint32_t, char*, uint64_t);
void ref_polisher_close(void*);
}

template <typename T>
static void put(FILE* f, const std::vector<T>& v) {
    fwrite(v.data(), sizeof(T), v.size(), f);
}

int main(int argc, char** argv) {
    if (argc != 14) {
        fprintf(stderr, "usage: %s reads overlaps target fragment w q e trim m x g threads out.bin\n", argv[0]);
        return 2;
    }
    void* h = ref_polisher_open(argv[1], argv[2], argv[3], atoi(argv[4]), static_cast<uint32_t>(atoi(argv[5])),
                                atof(argv[6]), atof(argv[7]), atoi(argv[8]), static_cast<int8_t>(atoi(argv[9])),
                                static_cast<int8_t>(atoi(argv[10])), static_cast<int8_t>(atoi(argv[11])),
                                static_cast<uint32_t>(atoi(argv[12])));
    uint64_t c[3];
    ref_polisher_counts(h, c);
    std::vector<char> bases(c[2]), quals(c[2]);
    std::vector<uint64_t> seq_off(c[1] + 1), win_target(c[0]);
    std::vector<uint8_t> has_q(c[1]), win_type(c[0]);
    std::vector<uint32_t> beg(c[1]), end(c[1]), first(c[0] + 1), win_rank(c[0]);
    ref_polisher_export(h, bases.data(), quals.data(), seq_off.data(), has_q.data(), beg.data(), end.data(),
                        first.data(), win_type.data(), win_target.data(), win_rank.data());
    uint32_t npol = ref_polisher_polish(h);
    FILE* f = fopen(argv[13], "wb");
    if (!f) return 3;
    uint64_t hdr[4] = {c[0], c[1], c[2], npol};
    fwrite(hdr, 8, 4, f);
    put(f, bases); put(f, quals); put(f, seq_off); put(f, has_q); put(f, beg); put(f, end); put(f, first);
    put(f, win_type); put(f, win_target); put(f, win_rank);
    std::vector<char> buf(1 << 20);
    for (uint32_t w = 0; w < c[0]; ++w) {
        uint32_t n = ref_polisher_window_consensus(h, w, buf.data(), static_cast<uint32_t>(buf.size()));
        fwrite(&n, 4, 1, f);
        fwrite(buf.data(), 1, n, f);
    }
    std::vector<char> data(1 << 27), name(4096);
    for (uint32_t i = 0; i < npol; ++i) {
        uint64_t n = ref_polisher_polished(h, i, name.data(), 4096, data.data(), data.size());
        uint32_t nl = static_cast<uint32_t>(std::string(name.data()).size());
        fwrite(&nl, 4, 1, f);
        fwrite(name.data(), 1, nl, f);
        fwrite(&n, 8, 1, f);
        fwrite(data.data(), 1, n, f);
    }
    fclose(f);
    ref_polisher_close(h);
    return 0;
}
