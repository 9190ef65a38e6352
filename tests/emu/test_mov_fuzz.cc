// tests/emu/test_mov_fuzz.cc -- TEST INFRASTRUCTURE ONLY.
// The QuickTime reader of hap_b200/csrc/hap_mov.h under AddressSanitizer / UBSan on damaged movies: a valid file
// (written by the writer in the same header) is truncated, bit-flipped and spliced a few thousand times; the reader
// must either refuse the file or hand out frames that lie inside it -- never crash, never read out of bounds.
#include "hap_mov.h"

#include <random>
#include <stdlib.h>
#include <unistd.h>

static std::vector<uint8_t> slurp(const char *path)
{
    std::vector<uint8_t> v;
    FILE *f = fopen(path, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize((size_t)n);
    if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

static void spill(const char *path, const std::vector<uint8_t> &v)
{
    FILE *f = fopen(path, "wb");
    if (v.size()) fwrite(v.data(), 1, v.size(), f);
    fclose(f);
}

int main(int argc, char **argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 3000;
    char good[] = "/tmp/hapmov_fuzz_good_XXXXXX", bad[] = "/tmp/hapmov_fuzz_bad_XXXXXX";
    close(mkstemp(good));
    close(mkstemp(bad));
    std::mt19937 rng(20260923);
    // a valid movie: 9 frames of different sizes
    {
        HapB200Mov *w = hapmov::create(good, HAPB200_FOURCC('H', 'a', 'p', 'Y'), 64, 32, 600);
        if (!w) return 2;
        for (int i = 0; i < 9; i++) {
            std::vector<uint8_t> fr(100 + 37 * i);
            for (auto &b : fr) b = (uint8_t)rng();
            if (hapmov::write_frame(w, fr.data(), fr.size(), 20 + i) != HapResult_No_Error) return 2;
        }
        if (hapmov::finish(w) != HapResult_No_Error) return 2;
        fclose(w->f);
        delete w;
    }
    const std::vector<uint8_t> base = slurp(good);
    if (base.empty()) return 2;
    int opened = 0, refused = 0;
    std::vector<uint8_t> buf(1 << 16);
    for (int r = 0; r < rounds; r++) {
        std::vector<uint8_t> v = base;
        const int what = r % 4;
        if (what == 0) {
            v.resize(rng() % v.size());
        } else if (what == 1) {
            for (int k = 0, n = 1 + rng() % 4; k < n; k++) v[rng() % v.size()] ^= (uint8_t)(1u << (rng() % 8));
        } else if (what == 2) {
            // overwrite a 32-bit field (sizes, counts, offsets live in them) with an extreme value
            const size_t at = rng() % (v.size() - 4);
            const uint32_t vals[6] = {0, 1, 0x7FFFFFFFu, 0xFFFFFFFFu, (uint32_t)v.size(), (uint32_t)rng()};
            const uint32_t x = vals[rng() % 6];
            v[at] = (uint8_t)(x >> 24); v[at + 1] = (uint8_t)(x >> 16); v[at + 2] = (uint8_t)(x >> 8); v[at + 3] = (uint8_t)x;
        } else {
            // splice: copy one region over another
            const size_t n = 1 + rng() % 64, a = rng() % (v.size() - n), b = rng() % (v.size() - n);
            memmove(&v[a], &v[b], n);
        }
        spill(bad, v);
        HapB200Mov *m = hapmov::open_read(bad);
        if (!m) { refused++; continue; }
        opened++;
        for (size_t i = 0; i < m->size.size(); i++) {
            // what HapB200MovReadFrame does
            if (m->size[i] > buf.size()) continue;
            if (fseeko(m->f, (off_t)m->offset[i], SEEK_SET) != 0 || fread(buf.data(), 1, m->size[i], m->f) != m->size[i]) {
                fprintf(stderr, "round %d: frame %zu of an accepted movie lies outside the file\n", r, i);
                return 1;
            }
        }
        fclose(m->f);
        delete m;
    }
    unlink(good);
    unlink(bad);
    printf("%d damaged movies: %d refused, %d opened and read without a fault\n", rounds, refused, opened);
    return 0;
}
