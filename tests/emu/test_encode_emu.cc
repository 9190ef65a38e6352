// tests/emu/test_encode_emu.cc -- TEST INFRASTRUCTURE ONLY.
// Runs the CUDA encode path (K5 fragment compressor, K6 plan + place) under the fiber SIMT emulator
// and checks the produced frames with the CPU oracle decoder (and the unmodified reference, when
// oracle/_ref/libhap_ref.so is present): payload bit-exact, container identical where it must be.
#define HAPB200_EMU
#include "hap_assemble.cuh"
#include "hap_host.h"
#include "snappy_decode.cuh"
#include "decode_emu.h"

extern "C" {
#include "bc_oracle.h"
#include "hap_oracle.h"
#include "snappy_oracle.h"
}
#include <dlfcn.h>
#include <random>
#include <string>

using namespace hapb200;
static int g_fail = 0;
typedef unsigned (*decode_fn)(const void *, unsigned long, unsigned, orc_decode_cb, void *, void *, unsigned long,
                              unsigned long *, unsigned *);
static decode_fn g_ref_decode = nullptr;

static void serial_cb(orc_work_fn fn, void *p, unsigned n, void *) { for (unsigned i = 0; i < n; i++) fn(p, i); }

struct Tex { std::vector<uint8_t> data; unsigned fmt, compressor, chunks; };

static std::vector<uint8_t> gpu_encode(const std::vector<Tex> &tex, int mode, uint32_t write_index)
{
    TextureArgs ta[2];
    for (size_t i = 0; i < tex.size(); i++) ta[i] = TextureArgs{tex[i].data.size(), tex[i].fmt, tex[i].compressor, tex[i].chunks};
    FrameGeom G;
    if (build_frame_geom((uint32_t)tex.size(), ta, G) != 0) abort();
    // one device-style input buffer holding both textures, 16-byte aligned offsets, exact size + no slack
    size_t off1 = (tex[0].data.size() + 15) & ~(size_t)15;
    std::vector<uint8_t> in(off1 + (tex.size() > 1 ? tex[1].data.size() : 0));
    memcpy(in.data(), tex[0].data.data(), tex[0].data.size());
    if (tex.size() > 1) { memcpy(in.data() + off1, tex[1].data.data(), tex[1].data.size()); G.s[1].in_offset = off1; }
    unsigned long lens[2], cap;
    unsigned fmts[2], chunks[2];
    for (size_t i = 0; i < tex.size(); i++) { lens[i] = tex[i].data.size(); fmts[i] = tex[i].fmt; chunks[i] = tex[i].chunks; }
    cap = orc_HapMaxEncodedLength((unsigned)tex.size(), lens, fmts, chunks);
    std::vector<uint8_t> scratch((size_t)G.frags_per_frame * kFragCap), out(cap, 0xEE);
    std::vector<uint32_t> fsize(G.frags_per_frame), fdst(G.frags_per_frame), fidx(G.frags_per_frame);
    std::vector<uint8_t> fent((size_t)G.frags_per_frame * kFragEntryStride, 0xDD);
    unsigned long long used = 0;
    emu::g_order_mode() = mode;
    // a grid smaller than the fragment count, as on the device: CTAs stride over the fragments (modes alternate 1, 2, 3 CTAs)
    const unsigned k5_grid = G.frags_per_frame < (unsigned)(1 + mode % 3) ? G.frags_per_frame : (unsigned)(1 + mode % 3);
    HAP_LAUNCH(snappy_encode_fragments_kernel, dim3(k5_grid), dim3(kEncThreads), sizeof(EncodeSmem), nullptr,
               in.data(), G, (uint32_t)G.frags_per_frame, scratch.data(), fsize.data(), (write_index & 1) ? fent.data() : (uint8_t *)nullptr);
    // write_index: bit 0 = fragment index, bit 1 = chunk offset table with aligned chunk starts
    HAP_LAUNCH(hap_plan_frames_kernel, dim3(1), dim3(kPlanThreads), 0, nullptr, G, in.data(), fsize.data(), fdst.data(), fidx.data(),
               write_index, out.data(), (uint64_t)cap, &used);
    HAP_LAUNCH(hap_place_fragments_kernel, dim3(G.frags_per_frame), dim3(kPlaceThreads), 0, nullptr, G, in.data(),
               scratch.data(), fsize.data(), fdst.data(), fidx.data(), (write_index & 1) ? fent.data() : (const uint8_t *)nullptr, out.data(),
               (uint64_t)cap);
    if (used > cap) { fprintf(stderr, "used %llu > cap %lu\n", used, cap); abort(); }
    out.resize(used);
    return out;
}

static void check(const std::string &name, const std::vector<Tex> &tex, int mode, uint32_t write_index, double *ratio = nullptr)
{
    if (getenv("EMU_TRACE")) fprintf(stderr, "case %s\n", name.c_str());
    std::vector<uint8_t> frame = gpu_encode(tex, mode, write_index);
    bool ok = true;
    std::string why;
    unsigned cnt = 0;
    if (orc_HapGetFrameTextureCount(frame.data(), frame.size(), &cnt) != 0 || cnt != tex.size()) { ok = false; why = "texture count"; }
    // oracle encode of the same input: container decisions that do not depend on Snappy bytes must agree
    {
        const void *ins[2]; unsigned long lens[2]; unsigned fmts[2], comps[2], chunks[2];
        for (size_t i = 0; i < tex.size(); i++) { ins[i] = tex[i].data.data(); lens[i] = tex[i].data.size(); fmts[i] = tex[i].fmt; comps[i] = tex[i].compressor; chunks[i] = tex[i].chunks; }
        unsigned long cap = orc_HapMaxEncodedLength((unsigned)tex.size(), lens, fmts, chunks), oused = 0;
        std::vector<uint8_t> of(cap);
        unsigned r = orc_HapEncode((unsigned)tex.size(), ins, lens, fmts, comps, chunks, of.data(), cap, &oused);
        if (r != 0) { ok = false; why = "oracle encode failed"; }
        bool all_none = true;
        for (auto &t : tex) all_none = all_none && t.compressor == 0;
        if (all_none && (oused != frame.size() || memcmp(of.data(), frame.data(), oused) != 0)) { ok = false; why = "None frame differs from oracle"; }
        for (unsigned i = 0; i < tex.size() && ok; i++) {
            int kc_o = 0, kc_g = 0;
            orc_HapGetFrameTextureChunkCount(of.data(), oused, i, &kc_o);
            orc_HapGetFrameTextureChunkCount(frame.data(), frame.size(), i, &kc_g);
            // the raw/complex decision may differ when sizes are borderline; chunk counts agree when both are complex
            if (kc_o != kc_g && kc_o != 1 && kc_g != 1) { ok = false; why = "chunk count differs from oracle"; }
        }
        if (ratio) *ratio = (double)frame.size() / (double)oused;
    }
    for (unsigned i = 0; i < tex.size() && ok; i++) {
        for (int which = 0; which < 2; which++) {
            decode_fn fn = which == 0 ? (decode_fn)orc_HapDecode : g_ref_decode;
            if (!fn) continue;
            std::vector<uint8_t> back(tex[i].data.size() + 8, 0x77);
            unsigned long used = 0; unsigned fmt = 0;
            unsigned r = fn(frame.data(), frame.size(), i, serial_cb, nullptr, back.data(), tex[i].data.size(), &used, &fmt);
            size_t expect = tex[i].data.size();
            if (r != 0 || fmt != tex[i].fmt) { ok = false; why = std::string(which ? "reference" : "oracle") + " decode result " + std::to_string(r); break; }
            // on the Snappy path bytes/chunks truncates, trailing bytes are not carried
            if (used != expect) {
                uint32_t k = limited_chunk_count(expect, tex[i].fmt, tex[i].chunks);
                if (used != (expect / k) * k) { ok = false; why = "decoded length"; break; }
            }
            if (memcmp(back.data(), tex[i].data.data(), used) != 0) { ok = false; why = std::string(which ? "reference" : "oracle") + " payload mismatch"; break; }
        }
    }
    // and through our own decode kernels (K7), jobs built with the host-side parser of the product: with the embedded index
    // (when the frame carries one), without it, and with a damaged index (the damaged chunks must be decoded again
    // from their streams alone and still give the payload)
    for (int variant = 0; variant < 3 && ok; variant++) {
        std::vector<uint8_t> fr = frame;
        for (unsigned i = 0; i < tex.size() && ok; i++) {
            Located loc;
            if (locate_texture(fr.data(), (uint32_t)fr.size(), i, loc) != 0) { ok = false; why = "locate"; break; }
            uint8_t *sec = fr.data() + loc.offset;
            std::vector<ChunkJob> jobs;
            std::vector<uint8_t> back(tex[i].data.size() + 64, 0x99);
            bool has_index = false;
            uint64_t in_sum = 0;
            if (((loc.type >> 4) & 0xF) == kHapComplex) {
                ChunkTables t; t.count = 0;
                if (parse_decode_instructions(sec, loc.len, t) != 0) { ok = false; why = "parse DI"; break; }
                FragmentIndex ix;
                const bool have_ix = locate_fragment_index(fr.data(), (uint32_t)fr.size(), ix);
                const uint32_t tables = kIndexHeaderBytes + 4u * (have_ix ? ix.chunks[0] + ix.chunks[1] : 0u);
                if (variant == 2 && have_ix && i == 0 && ix.len > tables + 8) {
                    // damage a few bytes of the records (sizes or entries), never the header
                    std::mt19937 r2(name.size() * 131 + mode);
                    for (int q = 0; q < 3; q++) fr[ix.body + tables + r2() % (ix.len - tables)] ^= (uint8_t)(1u << (r2() % 8));
                }
                uint64_t in_run = 0, out_run = 0;
                for (int c = 0; c < t.count; c++) {
                    uint32_t cc = sec[t.compressors + c], sz = rd_le32(sec + t.sizes + 4 * c), usz = sz;
                    ChunkJob j;
                    const uint64_t at = t.offsets != 0xFFFFFFFFu ? rd_le32(sec + t.offsets + 4 * c) : in_run;
                    if (cc == kHapChunkSnappy && !snappy_preamble(sec + t.data + at, sz, usz)) { ok = false; why = "preamble"; break; }
                    j.src = sec + t.data + at; j.dst = back.data() + 32 + out_run; j.src_bytes = sz; j.dst_bytes = usz; j.compressor = cc;
                    j.status = 99; j.index = nullptr; j.index_bytes = 0; j.mode = kJobUndecided; j.win_base = 0; j.win_count = 0;
                    uint32_t ioff = 0, ib = 0;
                    if (cc == kHapChunkSnappy && have_ix && fragment_index_record(fr.data(), ix, i, (uint32_t)t.count, (uint32_t)c, ioff, ib)) { j.index = fr.data() + ioff; j.index_bytes = ib; has_index = true; }
                    jobs.push_back(j);
                    in_run += sz; out_run += usz; in_sum += sz;
                }
                if ((write_index & 2) && t.offsets == 0xFFFFFFFFu) { ok = false; why = "offset table missing"; break; }
                if ((write_index & 2)) {
                    // every chunk starts on a 16-byte boundary of the frame, gaps are zero
                    for (int c = 0; c < t.count && ok; c++) {
                        const uint32_t o = rd_le32(sec + t.offsets + 4 * c);
                        if (((loc.offset + t.data + o) & 15) != 0) { ok = false; why = "chunk start not aligned"; }
                    }
                }
                if ((write_index & 1) && !has_index) {
                    bool any_snappy = false;
                    for (auto &j : jobs) any_snappy = any_snappy || j.compressor == kHapChunkSnappy;
                    if (any_snappy) { ok = false; why = "index section missing"; break; }
                }
                if (!(write_index & 1) && have_ix) { ok = false; why = "unexpected index section"; break; }
            } else {
                ChunkJob j;
                j.src = sec; j.dst = back.data() + 32; j.src_bytes = loc.len; j.dst_bytes = loc.len; j.compressor = kHapChunkRaw;
                j.status = 99; j.index = nullptr; j.index_bytes = 0; j.mode = kJobUndecided; j.win_base = 0; j.win_count = 0;
                jobs.push_back(j);
                in_sum = loc.len;
            }
            if (!ok) break;
            if (variant > 0 && !has_index) continue;   // nothing new to try
            uint32_t wo[2] = {0, 0};
            if (getenv("EMU_TRACE")) fprintf(stderr, "  decode variant %d texture %u jobs %zu has_index %d\n", variant, i, jobs.size(), (int)has_index);
            decode_jobs_emu(jobs.data(), (uint32_t)jobs.size(), in_sum, tex[i].data.size(), variant == 1 ? 0u : 1u, 1 + mode % 3, wo);
            size_t total = 0;
            for (auto &j : jobs) { if (j.status != 0) { ok = false; why = "K7 status " + std::to_string(j.status) + " variant " + std::to_string(variant); } total += j.dst_bytes; }
            if (ok && memcmp(back.data() + 32, tex[i].data.data(), total) != 0) { ok = false; why = "K7 payload mismatch, variant " + std::to_string(variant); }
            for (int g = 0; g < 32 && ok; g++) if (back[g] != 0x99 || back[32 + total + g] != 0x99) { ok = false; why = "K7 wrote outside its chunk"; }
            if (ok && variant == 0 && has_index && wo[1] != 0) { ok = false; why = "a sound embedded index was rejected"; }
        }
    }
    if (!ok) { g_fail++; fprintf(stderr, "FAIL %s mode %d: %s (frame %zu bytes)\n", name.c_str(), mode, why.c_str(), frame.size()); }
}

int main(int argc, char **argv)
{
    int modes = argc > 1 ? atoi(argv[1]) : 3;
    const char *refso = argc > 2 ? argv[2] : nullptr;
    if (refso) {
        void *h = dlopen(refso, RTLD_NOW | RTLD_LOCAL);
        if (h) g_ref_decode = (decode_fn)dlsym(h, "HapDecode");
        printf("reference decoder: %s\n", g_ref_decode ? "loaded" : "unavailable");
    }
    std::mt19937 rng(777);
    // a 256x128 picture with gradients + noise + bars, block-compressed by the oracle
    const int W = 256, H = 128;
    std::vector<uint8_t> img(W * H * 4);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint8_t *p = &img[4 * (y * W + x)];
            bool bar = y < 16 || y >= H - 16;
            for (int c = 0; c < 3; c++) p[c] = bar ? 16 : (uint8_t)(40 + (x * (c + 1)) % 160 + (y * 3) % 40 + (rng() % 4));
            p[3] = (uint8_t)(x);
        }
    auto blocks = [&](int kind) {
        std::vector<uint8_t> b((W / 4) * (H / 4) * (kind == 0 || kind == 3 ? 8 : 16));
        if (kind == 0) orc_bc1_encode_clusterfit(img.data(), W, H, b.data(), 1);
        else if (kind == 1) orc_bc3_encode_clusterfit(img.data(), W, H, b.data(), 1);
        else if (kind == 2) orc_ycocg_dxt5_encode_clusterfit(img.data(), W, H, b.data(), 1);
        else orc_bc4_encode_squish(img.data(), W, H, 3, b.data());
        return b;
    };
    std::vector<uint8_t> dxt1 = blocks(0), dxt5 = blocks(1), ycocg = blocks(2), rgtc = blocks(3);
    std::vector<uint8_t> noise(70000 - 70000 % 16), flat(16 * 5000);
    for (auto &b : noise) b = (uint8_t)rng();
    for (size_t i = 0; i < flat.size(); i++) flat[i] = (uint8_t)(0x30 + (i % 16));
    // bigger than one fragment, so chunk streams are concatenations of fragment streams
    std::vector<uint8_t> big;
    for (int r = 0; r < 5; r++) big.insert(big.end(), ycocg.begin(), ycocg.end());
    for (size_t i = 0; i < big.size(); i += 977) big[i] ^= (uint8_t)rng();

    struct Case { std::string name; std::vector<Tex> tex; };
    std::vector<Case> cases = {
        {"dxt1_1chunk", {{dxt1, 0x83F0, 1, 1}}},
        {"dxt1_3chunks", {{dxt1, 0x83F0, 1, 3}}},
        {"dxt5_4chunks", {{dxt5, 0x83F3, 1, 4}}},
        {"ycocg_8chunks", {{ycocg, 0x01, 1, 8}}},
        {"ycocg_limit_7_to_x", {{ycocg, 0x01, 1, 7}}},
        {"rgtc_2chunks", {{rgtc, 0x8DBB, 1, 2}}},
        {"hapm", {{ycocg, 0x01, 1, 4}, {rgtc, 0x8DBB, 1, 2}}},
        {"hapm_mixed_none", {{ycocg, 0x01, 0, 4}, {rgtc, 0x8DBB, 1, 2}}},
        {"none_single", {{dxt5, 0x83F3, 0, 5}}},
        {"none_pair", {{ycocg, 0x01, 0, 1}, {rgtc, 0x8DBB, 0, 1}}},
        {"noise_fallback_whole", {{noise, 0x83F3, 1, 2}}},
        {"flat_rle", {{flat, 0x01, 1, 2}}},
        {"big_multi_fragment", {{big, 0x01, 1, 2}}},
        {"big_1chunk", {{big, 0x8E8C, 1, 1}}},
        {"tiny_16", {{std::vector<uint8_t>(16, 0x55), 0x01, 1, 1}}},
        {"tiny_8", {{std::vector<uint8_t>(8, 0x11), 0x83F0, 1, 1}}},
        {"kat_a_64x55", {{std::vector<uint8_t>(64, 0x55), 0x83F0, 1, 1}}},
    };
    // per-chunk fallback: half compressible, half noise, two chunks
    {
        std::vector<uint8_t> mix(ycocg.begin(), ycocg.begin() + 16384);
        mix.insert(mix.end(), noise.begin(), noise.begin() + 16384);
        cases.push_back({"mixed_chunk_fallback", {{mix, 0x01, 1, 2}}});
    }
    // random shapes: sizes that are not multiples of the fragment or of the chunk count, 1..9 chunks, content
    // stitched from picture blocks, repeated blocks, flat runs and noise (exercises partial fragments, chunk-count
    // limiting, both fallbacks and the persistent grid's striding in combinations no hand-made case has)
    for (int r = 0; r < 8; r++) {
        const size_t blocks16 = 200 + rng() % 12000;
        std::vector<uint8_t> v;
        while (v.size() < blocks16 * 16) {
            const int what = rng() % 5;
            const size_t n = 16 * (1 + rng() % 700);
            if (what == 0) { size_t at = (rng() % (ycocg.size() / 16 - n / 16 - 1)) * 16; v.insert(v.end(), ycocg.begin() + at, ycocg.begin() + at + n); }
            else if (what == 1 && v.size() >= 4096) { size_t at = v.size() - 16 * (1 + rng() % 200); for (size_t i = 0; i < n; i++) v.push_back(v[at + i % 16]); }
            else if (what == 2) { uint8_t b16[16]; for (auto &b : b16) b = (uint8_t)rng(); for (size_t i = 0; i < n; i++) v.push_back(b16[i % 16]); }
            else if (what == 3) { for (size_t i = 0; i < n; i++) v.push_back((uint8_t)rng()); }
            else { for (size_t i = 0; i < n; i++) v.push_back((uint8_t)((i / 16) & 0xFF)); }
        }
        v.resize(blocks16 * 16);
        cases.push_back({"random_" + std::to_string(r), {{v, 0x01, 1, (unsigned)(1 + rng() % 9)}}});
    }
    for (int mode = 0; mode < modes; mode++)
        for (auto &c : cases) {
            double ratio = 0, ratio_i = 0;
            check(c.name, c.tex, mode, 0, &ratio);
            check(c.name + "+index", c.tex, mode, 1, &ratio_i);
            if (mode == 0) {   // (the chunk offset table changes host-side arithmetic only: one thread order is enough)
                double ratio_o = 0;
                check(c.name + "+offsets", c.tex, mode, 2, &ratio_o);
                check(c.name + "+index+offsets", c.tex, mode, 3, &ratio_o);
            }
            if (mode == 0) printf("  %-26s size vs oracle-encoded frame: %.3f   with fragment index: %.3f\n", c.name.c_str(), ratio, ratio_i);
        }
    g_fail += g_decode_emu_overflow;
    printf("%zu cases x %d modes, %d failures\n", cases.size(), modes, g_fail);
    return g_fail ? 1 : 0;
}
