// hot_kernels.cu -- the PERLND / IMPLND kernels instantiated on a batch-wide option word (StaticFlags).  Their own translation
// unit because they are compiled with the IEEE division sequence INLINE: with the sections a known option word switches off
// gone, the kernel is small enough that ~45 inline copies cost less than 18 warp-level calls per melt-season interval (each call
// and return is a non-sequential fetch: stall_no_inst), +1.6 % over the year.  The general kernels (perlnd_kernel.cu, implnd_kernel.cu) keep the
// single out-of-line copy, which is worth 16 % on the full chain.
#define HSP_INLINE_DIV
#include "io_configs.h"
#include "land_kernels.cuh"
#include "launch.h"

#ifndef HSP_MAXREG_HOT
#define HSP_MAXREG_HOT 128
#endif

template <int ACT_CT, class IO, class FW, int MAXREG = HSP_MAXREG_HOT, bool SHARED = false>
static cudaError_t launch_hot(LandLaunch &L, int sub_pref, cudaStream_t st, const char **name, const char *nm, int *grid_out) {
    const size_t smem = land_smem_bytes(IO::NF, perlnd_hot(ACT_CT), perlnd_ncold<true, IO, ACT_CT>(), SHARED);
    auto k = perlnd_kernel<ACT_CT, true, IO, MAXREG, SHARED, FW>;
    int cap = 0;
    cudaError_t e = hsp2g_prepare_kernel((const void *)k, smem, LAND_BLOCK, &cap);
    if (e != cudaSuccess) return e;
    hsp2g_plan_tasks(L, sub_pref, cap, LAND_WARPS, grid_out);
    HSP2G_LAUNCH(k, (unsigned)*grid_out, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

// true when a specialised kernel exists for this launch (and was launched; *err holds the result)
bool hsp2g_launch_perlnd_hot(LandLaunch &L, bool hourly, int sub_pref, cudaStream_t st, const char **name, int *grid_out, cudaError_t *err) {
    constexpr int SP = (int)(HSP2G_ACT_SNOW | HSP2G_ACT_PWATER);
    if (!hourly || L.act != (unsigned)SP || !L.flags_uniform) return false;
    // a replicated ensemble: every segment carries test10 PERLND 1's option word (test10.uci:65-169)
    if (L.flags_word != FLAGW_TEST10_P001) return false;
    typedef StaticFlags<FLAGW_TEST10_P001> FW;
    if (L.met) {  // shared-met mode (hsp2g_set_shared_forcing)
        if (!io_matches<IO_SnowPwaterDefault>(L)) return false;
        *err = launch_hot<SP, IO_SnowPwaterDefault, FW, HSP_MAXREG_HOT, true>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=default24,met=shared,options=test10-P001>", grid_out);
        return true;
    }
    if (io_matches<IO_SnowPwaterDefault>(L))
        *err = launch_hot<SP, IO_SnowPwaterDefault, FW>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=default24,options=test10-P001>", grid_out);
    else if (io_matches<IO_SnowPwaterFlows>(L))
        *err = launch_hot<SP, IO_SnowPwaterFlows, FW>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=flows3,options=test10-P001>", grid_out);
    else if (io_matches<IO_SnowPwaterAll>(L))
        *err = launch_hot<SP, IO_SnowPwaterAll, FW, 168>(L, sub_pref, st, name, "perlnd_kernel<SNOW|PWATER,hourly,save=all53,options=test10-P001>", grid_out);
    else
        return false;
    return true;
}

template <int ACT_CT, class IO, class FW>
static cudaError_t launch_hot_implnd(LandLaunch &L, int sub_pref, cudaStream_t st, const char **name, const char *nm, int *grid_out) {
    const size_t smem = land_smem_bytes(IO::NF, IMPLND_HOT, implnd_ncold<true, IO>(), false);
    auto k = implnd_kernel<ACT_CT, true, IO, 128, false, FW>;
    int cap = 0;
    cudaError_t e = hsp2g_prepare_kernel((const void *)k, smem, LAND_BLOCK, &cap);
    if (e != cudaSuccess) return e;
    hsp2g_plan_tasks(L, sub_pref, cap, LAND_WARPS, grid_out);
    HSP2G_LAUNCH(k, (unsigned)*grid_out, LAND_BLOCK, smem, st, L);
    if (name) *name = nm;
    return cudaGetLastError();
}

bool hsp2g_launch_implnd_hot(LandLaunch &L, bool hourly, int sub_pref, cudaStream_t st, const char **name, int *grid_out, cudaError_t *err) {
    constexpr int SI = (int)(HSP2G_ACT_SNOW | HSP2G_ACT_IWATER);
    constexpr int SIS = SI | (int)HSP2G_ACT_SOLIDS, SISG = SIS | (int)HSP2G_ACT_IWTGAS;
    if (L.met || !hourly || !L.flags_uniform) return false;
    // a replicated ensemble of test10's IMPLND 1 (test10.uci:194-248), with or without its SOLIDS / IWTGAS sections
    if (L.flags_word != FLAGW_TEST10_I001 || !io_matches<IO_SnowIwaterDefault>(L)) return false;
    typedef StaticFlags<FLAGW_TEST10_I001> FW;
    if (L.act == (unsigned)SI)
        *err = launch_hot_implnd<SI, IO_SnowIwaterDefault, FW>(L, sub_pref, st, name, "implnd_kernel<SNOW|IWATER,hourly,save=default16,options=test10-I001>", grid_out);
    else if (L.act == (unsigned)SIS)
        *err = launch_hot_implnd<SIS, IO_SnowIwaterDefault, FW>(L, sub_pref, st, name, "implnd_kernel<SNOW|IWATER|SOLIDS,hourly,save=default16,options=test10-I001>", grid_out);
    else if (L.act == (unsigned)SISG)
        *err = launch_hot_implnd<SISG, IO_SnowIwaterDefault, FW>(L, sub_pref, st, name, "implnd_kernel<SNOW|IWATER|SOLIDS|IWTGAS,hourly,save=default16,options=test10-I001>", grid_out);
    else
        return false;
    return true;
}
