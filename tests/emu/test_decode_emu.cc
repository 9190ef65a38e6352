// tests/emu/test_decode_emu.cc -- TEST INFRASTRUCTURE ONLY.
// Runs the CUDA Snappy decode kernel under the fiber SIMT emulator and checks it against the CPU
// oracle (oracle/snappy_oracle.c) on valid, pathological and corrupted streams.
#define HAPB200_EMU
#include "snappy_decode.cuh"
#include "decode_emu.h"

extern "C" {
#include "snappy_oracle.h"
}
#include <random>
#include <string>

using namespace hapb200;
static int g_fail = 0;

static std::vector<uint8_t> compress(const std::vector<uint8_t> &in)
{
    std::vector<uint8_t> out(orc_snappy_max_compressed_length(in.size()));
    size_t n = out.size();
    if (orc_snappy_compress(in.data(), in.size(), out.data(), &n) != ORC_SNAPPY_OK) abort();
    out.resize(n);
    return out;
}

static void check_stream(const std::string &name, const std::vector<uint8_t> &stream, int mode, int dmis = 0, int smis = 0)
{
    size_t want = 0;
    int pre = orc_snappy_uncompressed_length(stream.data(), stream.size(), &want);
    std::vector<uint8_t> ref(pre == ORC_SNAPPY_OK ? want + 1 : 1), got_store(ref.size() + 64 + 8, 0xCD);
    uint8_t *got = got_store.data() + dmis;  // destination misaligned by dmis bytes
    size_t rn = ref.size() - 1;
    int rs = pre == ORC_SNAPPY_OK ? orc_snappy_uncompress(stream.data(), stream.size(), ref.data(), &rn) : pre;
    ChunkJob job;
    std::vector<uint8_t> padded(stream.size() + 8);  // source misaligned by smis bytes
    memcpy(padded.data() + smis, stream.data(), stream.size());
    job.src = padded.data() + smis;
    job.dst = got + 32;
    job.src_bytes = (uint32_t)stream.size();
    job.dst_bytes = (uint32_t)(pre == ORC_SNAPPY_OK ? want : 0);
    job.compressor = kHapChunkSnappy;
    job.status = 99;
    job.index = nullptr;
    job.index_bytes = 0;
    job.mode = kJobUndecided;
    job.win_base = 0;
    job.win_count = 0;
    emu::g_order_mode() = mode;
    decode_jobs_emu(&job, 1, stream.size(), job.dst_bytes, 1);
    bool ok = true;
    if (rs == ORC_SNAPPY_OK) {
        if (job.status != HapResult_No_Error) ok = false;
        else if (memcmp(got + 32, ref.data(), want) != 0) ok = false;
    } else {
        if (job.status != HapResult_Bad_Frame) ok = false;
    }
    for (int i = 0; i < 32; i++)
        if (got[i] != 0xCD || got[32 + (pre == ORC_SNAPPY_OK ? want : 0) + i] != 0xCD) ok = false;  // wrote outside dst
    if (!ok) {
        g_fail++;
        size_t firstdiff = 0;
        if (rs == ORC_SNAPPY_OK)
            while (firstdiff < want && got[32 + firstdiff] == ref[firstdiff]) firstdiff++;
        fprintf(stderr, "FAIL %s mode %d: oracle %d kernel status %u len %zu firstdiff %zu\n", name.c_str(), mode, rs,
                job.status, want, firstdiff);
    }
}

int main(int argc, char **argv)
{
    int modes = argc > 1 ? atoi(argv[1]) : 3;
    std::mt19937 rng(12345);
    std::vector<std::pair<std::string, std::vector<uint8_t>>> cases;
    auto add = [&](const std::string &n, const std::vector<uint8_t> &raw) { cases.push_back({n, compress(raw)}); };

    add("empty", {});
    add("one", {42});
    add("zeros_300k", std::vector<uint8_t>(300000, 0));
    {
        std::vector<uint8_t> v(200000);
        for (auto &b : v) b = (uint8_t)rng();
        add("noise_200k", v);
    }
    {
        // DXT5-like: 16-byte blocks, many repeats of the block above (row stride 960 blocks) and neighbours
        std::vector<uint8_t> v(16 * 960 * 12);
        for (size_t b = 0; b < v.size() / 16; b++) {
            int mode = rng() % 8;
            for (int k = 0; k < 16; k++) {
                uint8_t x;
                if (mode < 3 && b >= 960) x = v[(b - 960) * 16 + k];
                else if (mode < 5 && b >= 1) x = (k < 8) ? v[(b - 1) * 16 + k] : (uint8_t)rng();
                else if (mode == 5 && b >= 1) x = v[(b - 1) * 16 + k];
                else x = (uint8_t)rng();
                v[b * 16 + k] = x;
            }
        }
        add("dxt5_like", v);
    }
    {
        // long literals interleaved with short matches (DXT1-like)
        std::vector<uint8_t> v(150000);
        for (auto &b : v) b = (uint8_t)rng();
        for (size_t i = 5000; i + 64 < v.size(); i += 3001) memcpy(&v[i], &v[i - 4096], 48);
        add("dxt1_like", v);
    }
    {
        // period-3 / period-7 runs and short periodic structures
        std::vector<uint8_t> v;
        for (int r = 0; r < 400; r++) {
            int period = 1 + rng() % 19, n = 10 + rng() % 900;
            std::vector<uint8_t> p(period);
            for (auto &b : p) b = (uint8_t)rng();
            for (int i = 0; i < n; i++) v.push_back(p[i % period]);
        }
        add("periodic_mix", v);
    }
    {
        // chains: block k copies half of block k-1 (dependency depth = run length)
        std::vector<uint8_t> v(16 * 3000);
        for (size_t b = 0; b < 3000; b++)
            for (int k = 0; k < 16; k++) v[b * 16 + k] = (k < 8 && b > 0) ? v[(b - 1) * 16 + k] : (uint8_t)rng();
        add("chain_depth", v);
    }
    {
        // dense stream: thousands of 4..7-byte copies out of a small dictionary -> far more elements per window than
        // descriptor slots, so windows are cut short and the adaptive span shrinks and grows again
        std::vector<uint8_t> dict(4096), v;
        for (auto &b : dict) b = (uint8_t)rng();
        v.insert(v.end(), dict.begin(), dict.end());
        for (int r = 0; r < 30000; r++) {
            int n = 4 + rng() % 4, at = rng() % (4096 - 8);
            for (int i = 0; i < n; i++) v.push_back(dict[at + i]);
        }
        for (int r = 0; r < 20000; r++) v.push_back((uint8_t)rng());  // then literal-heavy again
        add("dense_copies", v);
    }
    // hand-made stream with every element kind (same bytes as tests/golden "all_kinds")
    {
        std::vector<uint8_t> s;
        auto varint = [&](uint32_t n) { while (n >= 0x80) { s.push_back((n & 0x7F) | 0x80); n >>= 7; } s.push_back(n); };
        varint(20 + 7 + 64 + 10 + 100 + 300 + 33);
        s.push_back(19 << 2); for (int i = 1; i <= 20; i++) s.push_back(i);
        s.push_back(1 | (3 << 2)); s.push_back(20);
        s.push_back(2 | (63 << 2)); s.push_back(1); s.push_back(0);
        s.push_back(3 | (9 << 2)); s.push_back(30); s.push_back(0); s.push_back(0); s.push_back(0);
        s.push_back(60 << 2); s.push_back(99); for (int i = 0; i < 100; i++) s.push_back((i * 7) & 0xFF);
        s.push_back(61 << 2); s.push_back(299 & 0xFF); s.push_back(299 >> 8); for (int i = 0; i < 300; i++) s.push_back((i * 13 + 5) & 0xFF);
        s.push_back(2 | (32 << 2)); s.push_back(3); s.push_back(0);
        cases.push_back({"all_kinds", s});
    }
    // corrupted variants of a valid stream: truncations, bit flips, bad lengths
    {
        std::vector<uint8_t> base = cases[4].second;  // dxt5_like
        for (int k = 0; k < 24; k++) {
            std::vector<uint8_t> c = base;
            int what = k % 3;
            if (what == 0) c.resize(c.size() - 1 - rng() % 200);
            else if (what == 1) c[8 + rng() % (c.size() - 8)] ^= (uint8_t)(1u << (rng() % 8));
            else { size_t p = 3 + rng() % 64; c[p] = (uint8_t)rng(); c[p + 1] = (uint8_t)rng(); }
            cases.push_back({"corrupt_" + std::to_string(k), c});
        }
        std::vector<uint8_t> c = {8, 3 << 2, 'a', 'b', 'c', 'd', 2 | (3 << 2), 0, 0};
        cases.push_back({"offset_zero", c});
        c = {8, 3 << 2, 'a', 'b', 'c', 'd', 2 | (3 << 2), 5, 0};
        cases.push_back({"offset_far", c});
        c = {0x80, 0x80, 0x80, 0x80, 0x80};
        cases.push_back({"bad_varint", c});
    }
    for (int mode = 0; mode < modes; mode++)
        for (auto &c : cases) check_stream(c.first, c.second, mode);
    // every destination / source alignment (the kernel has word, funnel and byte paths)
    for (int dm = 0; dm < 4; dm++)
        for (int sm = 0; sm < 4; sm++)
            for (auto &c : cases) check_stream(c.first + "@d" + std::to_string(dm) + "s" + std::to_string(sm), c.second, 2, dm, sm);
    g_fail += g_decode_emu_overflow;
    printf("%zu cases x %d modes, %d failures, %llu barriers\n", cases.size(), modes, g_fail,
           (unsigned long long)emu::g_barriers());
    return g_fail ? 1 : 0;
}
