// land_kernels.cuh -- the fused per-segment time loops.  One thread = one land segment, one warp = one tile of
// 32 segments; a warp walks the chunk's intervals strictly in order (the recursion in time is sequential,
// SURVEY.md 5) with every storage in registers, takes its forcings from its TMA-filled shared-memory ring and
// streams the SAVEd series out time-major (land_common.cuh).
//
// Activity order inside an interval is the reference registry's (configuration.py:48-52):
//   PERLND: ATEMP -> SNOW -> PWATER -> SEDMNT -> PSTEMP -> PWTGAS        IMPLND: ATEMP -> SNOW -> IWATER
// All couplings are same-interval, so one pass per interval is exact.
//
// Template parameters
//   ACT_CT >= 0  fixes the activity mask at compile time (hot configurations); < 0 reads it from the launch record;
//   HOURLY       the run step is >= 60 min, where HRFG is identically 1 (calendar.py hoursval), which lets the
//                compiler drop most of SNOW's hour-gated carried scalars from the loop-carried register set;
//   IO           StaticIO<...> (forcing rows and SAVE set fixed at compile time) or DynIO (launch-record tables);
//   MAXREG       register cap per thread (__maxnreg__): resident warps per SM = 65536 / (32 * MAXREG rounded up to 8).
#pragma once
#include "iwater.cuh"
#include "iwtgas.cuh"
#include "land_common.cuh"
#include "pstemp.cuh"
#include "pwater.cuh"
#include "pwtgas.cuh"
#include "sedmnt.cuh"
#include "snow.cuh"
#include "solids.cuh"

#define IMPLND_HOT (HOT_SNOW | HOT_IWATER)
#define CHAIN_ACTS (HSP2G_ACT_ATEMP | HSP2G_ACT_SEDMNT | HSP2G_ACT_PSTEMP | HSP2G_ACT_PWTGAS)
// shared-memory plan of a PERLND kernel instantiation: hot-parameter set and cold-state rows depend on whether the
// sections beyond SNOW+PWATER can run (ACT_CT < 0: any mask)
__host__ __device__ constexpr bool perlnd_has_chain(int act_ct) { return act_ct < 0 || (act_ct & (int)CHAIN_ACTS) != 0; }
__host__ __device__ constexpr unsigned long long perlnd_hot(int act_ct) { return HOT_SNOW | HOT_PWATER | (perlnd_has_chain(act_ct) ? HOT_CHAIN : 0ull); }
// ... as a kernel with compile-time row sets stages them: a batch without lateral inflow series never reads the lateral factors,
// and the scour coefficients are read on intervals with surface runoff only (from global memory then, like RDCSN while it
// snows).  71 instead of 76 shared-memory rows per warp: SIX two-warp blocks of the full chain fit an SM (12 resident warps, what
// its 168 registers allow) instead of five (profiles/r3f_fullchain_all_plain_ncu_summary.md: shared memory was the limiter).
template <class IO>
__host__ __device__ constexpr unsigned long long perlnd_hot_io(int act_ct) {
    unsigned long long h = perlnd_hot(act_ct);
    if (IO::STATIC && perlnd_has_chain(act_ct)) {
        if (!((IO::FMASK >> F_SURLI) & 1ull)) h &= ~PBIT(SLIFAC);
        if (!((IO::FMASK >> F_IFWLI) & 1ull)) h &= ~PBIT(ILIFAC);
        if (!((IO::FMASK >> F_AGWLI) & 1ull)) h &= ~PBIT(ALIFAC);
        h &= ~(PBIT(KGER) | PBIT(JGER));
    }
    return h;
}
#define OPT_SED_RAIN_IS_PREC 1u  // SEDMNT: no RAINF in ts -> RAIN = PREC (SEDMNT.py:86-89)


// hourly kernels with a compile-time SAVE set keep SNOW's hour-gated scalars write-through (ColdW); the others in
// shared memory (ColdD)
template <bool W>
struct ColdPick {
    typedef ColdD type;
};
template <>
struct ColdPick<true> {
    typedef ColdW type;
};

template <class SEG>
__device__ __forceinline__ double atemp_step(const SEG &s, double prec) {
    // ATEMP.py:57-58 with k = delt*0.000833 (ATEMP.py:22)
    const double k = s.L.delt * 0.000833;
    const double lapse = prec > k ? 0.0035 : s.calp->lapse;
    const double airtmp = s.forcing(F_GATMP) - lapse * s.par(P_ELDAT);
    s.save(O_ATEMP_AIRTMP, airtmp);
    return airtmp;
}

// The time loop of a PERLND warp.  The __global__ wrappers below (and the ones hspsquared_b200/jit.py generates for one
// batch's exact configuration) only fix the template arguments.
template <int ACT_CT, bool HOURLY, class IO, bool SHARED, class FW, class LAY>
__device__ __forceinline__ void perlnd_body(const LandLaunch &L) {
    LAND_SMEM_DECL;
    int lane = threadIdx.x & (LAND_TILE - 1);
    const int wib = threadIdx.x / LAND_TILE;
    HSP_PIN32(lane);  // kept in a register: rematerialising it costs an S2R per use
    const unsigned act = ACT_CT >= 0 ? (unsigned)ACT_CT : L.act;
    const int nf = IO::STATIC ? IO::NF : L.nf;
    typedef Seg<IO, perlnd_hot_io<IO>(ACT_CT), SHARED, FW, LAY> SegT;
    typedef typename ColdPick<HOURLY && IO::STATIC>::type SNOWCOLD;
    constexpr int NCOLD = SnowStateT<SNOWCOLD>::SMEM_ROWS + NCOLD_PWATER + (perlnd_has_chain(ACT_CT) ? NCOLD_CHAIN : 0);
    PipeT<SHARED, LAY> pipe;
    pipe.init(L, smem, wib, lane, nf, SegT::NHOT + (SHARED ? nf : 0), NCOLD);
    Sched sched(L, lane, wib);
    int tile, t0, t1;
    while (sched.next(tile, t0, t1)) {  // one task = intervals [t0, t1) of one tile (land_common.cuh Sched)
        HSP_TASK_IDLE(sched, L)
        pipe.begin_task(tile, t0, t1);
        SegT s(L, tile, lane, t0, pipe.hot_area());

        SnowStateT<SNOWCOLD> sw;
        PwaterState pw;
        SedmntState sd;
        PstempState ps;
        double *cold = pipe.hot_area() + (SegT::NHOT + (SHARED ? nf : 0)) * LAND_TILE + lane;
        sw.bind(cold, s.sb);
        pw.bind(cold + SnowStateT<SNOWCOLD>::SMEM_ROWS * LAND_TILE, s.sb);
        PwtgasDaily pg;
        if (perlnd_has_chain(ACT_CT)) {
            ps.bind(cold + (SnowStateT<SNOWCOLD>::SMEM_ROWS + NCOLD_PWATER) * LAND_TILE);
            pg.bind(cold + (SnowStateT<SNOWCOLD>::SMEM_ROWS + NCOLD_PWATER + NCOLD_PSTEMP) * LAND_TILE);
        }
        if (act & HSP2G_ACT_SNOW) sw.load(s);
        if (act & HSP2G_ACT_PWATER) pw.load(s);
        if (act & HSP2G_ACT_SEDMNT) sd.load(s);
        if (act & HSP2G_ACT_PSTEMP) ps.load(s);

        const int t_lim = HSP_T_LIM(L, t0, t1);
        for (int t = t0; t < t_lim; ++t) {
            HSP_STEP_SYNC(t, t1)
            s.begin_step(pipe.acquire());
            const double prec = s.forcing(F_PREC);
            const double airtmp = (act & HSP2G_ACT_ATEMP) ? atemp_step(s, prec) : s.forcing(F_AIRTMP);

            SnowFlux sx;
            SnowLink link;
            if (act & HSP2G_ACT_SNOW) {
                snow_step<HOURLY>(s, sw, sx, airtmp, prec);
                link.rainf = sx.rainf;
                link.snocov = sw.snocov;
                link.wyield = sx.wyield;
                link.packi = sw.packi;
                if (s.metric()) {  // the water section reads the series SNOW stored, i.e. converted (SNOW.py:575,582)
                    link.wyield = link.wyield * 25.4;
                    link.packi = link.packi * 25.4;
                }
            } else {
                link.rainf = s.forcing(F_RAINF);
                link.snocov = s.forcing(F_SNOCOV);
                link.wyield = s.forcing(F_WYIELD);
                link.packi = s.forcing(F_PACKI);
            }
            if (s.metric()) link.wyield = link.wyield * 0.0394;  // ... and takes WYIELD back to inches (PWATER.py:244, IWATER.py:114)
            link.airtmp = airtmp;

            PwaterFlux fx;
            HSP_SECTION_SYNC();
            if (act & HSP2G_ACT_PWATER) {
                if (t == t0 || (s.calfl & HSP2G_CAL_NEWDAY)) pw.refresh_daily(s);
                pwater_step<HOURLY>(s, pw, link, prec, fx);
            } else {
                fx.suro = s.forcing(F_SURO);
                fx.surs = s.forcing(F_SURS);
                fx.ifwo = s.forcing(F_IFWO);
                fx.agwo = s.forcing(F_AGWO);
            }
            if (perlnd_has_chain(ACT_CT) && s.metric()) {
                // the later sections read the series the water section STORED, i.e. x * 25.4 (PWATER.py:667-682), and take them
                // back to inches with * 0.0394 (SEDMNT.py:115-116, PWTGAS.py:172-174); external series are in mm already
                if (act & HSP2G_ACT_PWATER) {
                    fx.suro = fx.suro * 25.4;
                    fx.surs = fx.surs * 25.4;
                    fx.ifwo = fx.ifwo * 25.4;
                    fx.agwo = fx.agwo * 25.4;
                }
                fx.suro = fx.suro * 0.0394;
                fx.surs = fx.surs * 0.0394;
                fx.ifwo = fx.ifwo * 0.0394;
                fx.agwo = fx.agwo * 0.0394;
            }
            if (act & HSP2G_ACT_SEDMNT) {
                const double rain = (s.opts() & OPT_SED_RAIN_IS_PREC) ? prec : link.rainf;
                sedmnt_step(s, sd, rain, prec, link.snocov, fx.suro, fx.surs);
            }
            SoilTemps soil;
            const bool newday = t == t0 || (s.calfl & HSP2G_CAL_NEWDAY);
            if (act & HSP2G_ACT_PSTEMP) {
                if (newday) ps.refresh_daily(s);
                if (L.first_chunk && t == 0) ps.airts = airtmp;  // PSTEMP.py:128
                pstemp_step<HOURLY>(s, ps, airtmp, soil);
            } else {
                soil.sltmp = s.forcing(F_SLTMP);
                soil.ultmp = s.forcing(F_ULTMP);
                soil.lgtmp = s.forcing(F_LGTMP);
                if (s.metric()) {  // external soil temperatures are deg C (PWTGAS.py:168-170)
                    soil.sltmp = (soil.sltmp * 9. / 5.) + 32.;
                    soil.ultmp = (soil.ultmp * 9. / 5.) + 32.;
                    soil.lgtmp = (soil.lgtmp * 9. / 5.) + 32.;
                }
            }
            if (act & HSP2G_ACT_PWTGAS) {
                if (newday) pg.refresh_daily(s);
                pwtgas_step(s, pg, soil, link.wyield, fx.suro, fx.ifwo, fx.agwo);
            }
            s.end_step();
            pipe.release(t);
        }
        if (act & HSP2G_ACT_SNOW) sw.store(s);
        if (act & HSP2G_ACT_PWATER) pw.store(s);
        if (act & HSP2G_ACT_SEDMNT) sd.store(s);
        if (act & HSP2G_ACT_PSTEMP) ps.store(s);
        pipe.end_task();
        sched.done(tile);
    }
}
template <int ACT_CT, bool HOURLY, class IO, int MAXREG, bool SHARED, class FW = DynFlags, class LAY = TiledLay>
__global__ void __launch_bounds__(LAND_BLOCK) HSP_MAXNREG(MAXREG) perlnd_kernel(const __grid_constant__ LandLaunch L) {
    perlnd_body<ACT_CT, HOURLY, IO, SHARED, FW, LAY>(L);
}

template <int ACT_CT, bool HOURLY, class IO, bool SHARED, class FW, class LAY>
__device__ __forceinline__ void implnd_body(const LandLaunch &L) {
    LAND_SMEM_DECL;
    int lane = threadIdx.x & (LAND_TILE - 1);
    const int wib = threadIdx.x / LAND_TILE;
    HSP_PIN32(lane);  // kept in a register: rematerialising it costs an S2R per use
    const unsigned act = ACT_CT >= 0 ? (unsigned)ACT_CT : L.act;
    const int nf = IO::STATIC ? IO::NF : L.nf;
    typedef Seg<IO, IMPLND_HOT, SHARED, FW, LAY> SegT;
    typedef typename ColdPick<HOURLY && IO::STATIC>::type SNOWCOLD;
    constexpr int NCOLD = SnowStateT<SNOWCOLD>::SMEM_ROWS;
    PipeT<SHARED, LAY> pipe;
    pipe.init(L, smem, wib, lane, nf, SegT::NHOT + (SHARED ? nf : 0), NCOLD);
    Sched sched(L, lane, wib);
    int tile, t0, t1;
    while (sched.next(tile, t0, t1)) {
        HSP_TASK_IDLE(sched, L)
        pipe.begin_task(tile, t0, t1);
        SegT s(L, tile, lane, t0, pipe.hot_area());

        SnowStateT<SNOWCOLD> sw;
        IwaterState iw;
        SolidsState sl;
        IwtgasState ig;
        sw.bind(pipe.hot_area() + (SegT::NHOT + (SHARED ? nf : 0)) * LAND_TILE + lane, s.sb);
        if (act & HSP2G_ACT_SNOW) sw.load(s);
        if (act & HSP2G_ACT_IWATER) iw.load(s);
        if (act & HSP2G_ACT_SOLIDS) sl.load(s);
        if (act & HSP2G_ACT_IWTGAS) ig.load(s);

        const int t_lim = HSP_T_LIM(L, t0, t1);
        for (int t = t0; t < t_lim; ++t) {
            HSP_STEP_SYNC(t, t1)
            s.begin_step(pipe.acquire());
            const double prec = s.forcing(F_PREC);
            const double airtmp = (act & HSP2G_ACT_ATEMP) ? atemp_step(s, prec) : s.forcing(F_AIRTMP);
            SnowFlux sx;
            SnowLink link;
            if (act & HSP2G_ACT_SNOW) {
                snow_step<HOURLY>(s, sw, sx, airtmp, prec);
                link.rainf = sx.rainf;
                link.snocov = sw.snocov;
                link.wyield = sx.wyield;
                link.packi = sw.packi;
                if (s.metric()) {  // the water section reads the series SNOW stored, i.e. converted (SNOW.py:575,582)
                    link.wyield = link.wyield * 25.4;
                    link.packi = link.packi * 25.4;
                }
            } else {
                link.rainf = s.forcing(F_RAINF);
                link.snocov = s.forcing(F_SNOCOV);
                link.wyield = s.forcing(F_WYIELD);
                link.packi = s.forcing(F_PACKI);
            }
            if (s.metric()) link.wyield = link.wyield * 0.0394;  // ... and takes WYIELD back to inches (PWATER.py:244, IWATER.py:114)
            link.airtmp = airtmp;
            double suro, surs;
            if (act & HSP2G_ACT_IWATER) {
                if (t == t0 || (s.calfl & HSP2G_CAL_NEWDAY)) iw.refresh_daily(s);
                iwater_step<HOURLY>(s, iw, link, prec, suro, surs);
            } else {
                suro = s.forcing(F_SURO);
                surs = s.forcing(F_SURS);
            }
            if ((act & (HSP2G_ACT_SOLIDS | HSP2G_ACT_IWTGAS)) && s.metric()) {
                // SOLIDS / IWTGAS read the series IWATER stored (x * 25.4, IWATER.py:281-282) back with * 0.0394 (SOLIDS.py:86-87,
                // IWTGAS.py:114); external series are in mm already
                if (act & HSP2G_ACT_IWATER) {
                    suro = suro * 25.4;
                    surs = surs * 25.4;
                }
                suro = suro * 0.0394;
                surs = surs * 0.0394;
            }
            if (act & HSP2G_ACT_SOLIDS) solids_step(s, sl, suro, surs, prec);
            if (act & HSP2G_ACT_IWTGAS) iwtgas_step(s, ig, airtmp, link.wyield, suro);
            s.end_step();
            pipe.release(t);
        }
        if (act & HSP2G_ACT_SNOW) sw.store(s);
        if (act & HSP2G_ACT_IWATER) iw.store(s);
        if (act & HSP2G_ACT_SOLIDS) sl.store(s);
        if (act & HSP2G_ACT_IWTGAS) ig.store(s);
        pipe.end_task();
        sched.done(tile);
    }
}
template <int ACT_CT, bool HOURLY, class IO, int MAXREG, bool SHARED, class FW = DynFlags, class LAY = TiledLay>
__global__ void __launch_bounds__(LAND_BLOCK) HSP_MAXNREG(MAXREG) implnd_kernel(const __grid_constant__ LandLaunch L) {
    implnd_body<ACT_CT, HOURLY, IO, SHARED, FW, LAY>(L);
}

// dynamic shared memory of one block: per warp PIPE_STAGES interval stages (forcing elements of fesz bytes) + the hot
// parameters + the cold state + one mbarrier per stage
__host__ __device__ constexpr size_t land_smem_bytes(int nf, unsigned long long hot, int ncold, bool shared, int fesz = 8) {
    return (size_t)LAND_WARPS * ((size_t)PIPE_STAGES * nf * LAND_TILE * (shared ? 8 : fesz) +
                                 (size_t)(ce_popc(hot) + (shared ? nf : 0) + ncold) * LAND_TILE * sizeof(double) + PIPE_STAGES * 8);
}
// cold-state rows of a kernel instantiation (must mirror the NCOLD constants inside the kernels)
template <bool HOURLY, class IO, int ACT_CT>
__host__ __device__ constexpr int perlnd_ncold() {
    return SnowStateT<typename ColdPick<HOURLY && IO::STATIC>::type>::SMEM_ROWS + NCOLD_PWATER +
           (perlnd_has_chain(ACT_CT) ? NCOLD_CHAIN : 0);
}
template <bool HOURLY, class IO>
__host__ __device__ constexpr int implnd_ncold() {
    return SnowStateT<typename ColdPick<HOURLY && IO::STATIC>::type>::SMEM_ROWS;
}
