// tests/emu/decode_emu.h -- TEST INFRASTRUCTURE ONLY: the device pipeline of hap_api.cu's launch_decode_jobs on the emulator
// (windows, on-the-fly index, execute, repair of chunks whose embedded index did not hold up).
#pragma once
#include "snappy_decode.cuh"
#include <vector>

static int g_decode_emu_overflow = 0;
static inline void decode_jobs_emu(hapb200::ChunkJob *jobs, uint32_t njobs, uint64_t in_bound, uint64_t out_bound, uint32_t use_index,
                                   unsigned ex_grid = 3, uint32_t *windows_out = nullptr)
{
    using namespace hapb200;
    const uint32_t entry_slots = (uint32_t)(2 * (in_bound / kIdxWin + njobs) + 16);
    const uint32_t win_cap = (uint32_t)(kIdxParts * (uint64_t)entry_slots + out_bound / kIndexFragBytes + 2ull * njobs + 16);
    std::vector<DecWin> wins(win_cap);
    std::vector<uint8_t> entries((size_t)entry_slots * kIdxThreads, 0xEE);
    std::vector<uint32_t> done(win_cap, 0);
    DecodeCtl ctl;
    memset(&ctl, 0, sizeof ctl);
    uint32_t any_left = 0;
    HAP_LAUNCH(hap_build_windows_kernel, dim3((njobs + 127) / 128), dim3(128), 0, nullptr, jobs, njobs, use_index, wins.data(), win_cap, &ctl);
    HAP_LAUNCH(snappy_index_kernel, dim3(njobs), dim3(kIdxThreads), sizeof(IndexSmem), nullptr, jobs, (int)njobs, 0u, wins.data(),
               win_cap, entries.data(), entry_slots, &ctl);
    HAP_LAUNCH(snappy_execute_kernel, dim3(ex_grid), dim3(kExThreads), sizeof(ExecSmem), nullptr, jobs, njobs, 0u, wins.data(), &ctl, done.data());
    if (use_index) {
        HAP_LAUNCH(hap_requeue_mismatched_kernel, dim3((njobs + 127) / 128), dim3(128), 0, nullptr, jobs, njobs, &any_left);
        HAP_LAUNCH(snappy_index_kernel, dim3(njobs), dim3(kIdxThreads), sizeof(IndexSmem), nullptr, jobs, (int)njobs, 1u, wins.data(),
                   win_cap, entries.data(), entry_slots, &ctl);
        HAP_LAUNCH(snappy_execute_kernel, dim3(ex_grid), dim3(kExThreads), sizeof(ExecSmem), nullptr, jobs, njobs, 1u, wins.data(), &ctl, done.data());
    }
    if (ctl.overflow) g_decode_emu_overflow++;
    if (windows_out) { windows_out[0] = ctl.n_windows; windows_out[1] = any_left; }
}
