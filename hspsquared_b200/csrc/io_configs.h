// io_configs.h -- the compile-time I/O configurations (StaticIO) the library ships besides the run-time one.
// A batch whose forcing rows and SAVE rows are in ascending id order and equal one of these sets runs the
// specialised kernel; anything else runs the DynIO kernel (same arithmetic, row tables read at run time).
#pragma once
#include "land_common.cuh"

constexpr unsigned long long fbit(int id) { return 1ull << id; }
constexpr unsigned long long olo(int oid) { return oid < 64 ? 1ull << oid : 0ull; }
constexpr unsigned long long ohi(int oid) { return oid >= 64 ? 1ull << (oid - 64) : 0ull; }

// the six meteorological series of a SNOW(+water) segment without ATEMP / CLOUD (BASELINE.json north_star)
constexpr unsigned long long FM_MET6 = fbit(F_PREC) | fbit(F_PETINP) | fbit(F_AIRTMP) | fbit(F_WINMOV) | fbit(F_SOLRAD) |
                                       fbit(F_DTMPG);
// ... with ATEMP computing AIRTMP from GATMP
constexpr unsigned long long FM_MET6_GATMP = (FM_MET6 & ~fbit(F_AIRTMP)) | fbit(F_GATMP);
constexpr unsigned long long FM_PREC_PET = fbit(F_PREC) | fbit(F_PETINP);

// default SAVE tables (hsp2tools/data/SaveTable.csv): SNOW 13, PWATER 11, IWATER 3
#define SAVE_SNOW_DEFAULT(B)                                                                                      \
    (B(O_SNOW_COVINX) | B(O_SNOW_DULL) | B(O_SNOW_PACKF) | B(O_SNOW_PACKI) | B(O_SNOW_PACKW) | B(O_SNOW_PACK) |   \
     B(O_SNOW_PAKTMP) | B(O_SNOW_RAINF) | B(O_SNOW_RDENPF) | B(O_SNOW_SKYCLR) | B(O_SNOW_SNOCOV) |                \
     B(O_SNOW_WYIELD) | B(O_SNOW_XLNMLT))
#define SAVE_PWATER_DEFAULT(B)                                                                                    \
    (B(O_PWATER_AGWO) | B(O_PWATER_AGWS) | B(O_PWATER_CEPS) | B(O_PWATER_GWVS) | B(O_PWATER_IFWO) |               \
     B(O_PWATER_IFWS) | B(O_PWATER_LZS) | B(O_PWATER_PERO) | B(O_PWATER_SURO) | B(O_PWATER_SURS) | B(O_PWATER_UZS))
#define SAVE_IWATER_DEFAULT(B) (B(O_IWATER_RETS) | B(O_IWATER_SURO) | B(O_IWATER_SURS))
// the three runoff components a CBP-style watershed run keeps (tests/land_spec/hwmA51800.uci:493-495)
#define SAVE_PWATER_FLOWS(B) (B(O_PWATER_SURO) | B(O_PWATER_IFWO) | B(O_PWATER_AGWO))

// all outputs of a section: ids are contiguous per section (hsp2g_ids.h)
constexpr unsigned long long range_lo(int a, int b) {  // bits [a, b) that fall in 0..63
    unsigned long long m = 0;
    for (int i = a; i < b; ++i) m |= olo(i);
    return m;
}
constexpr unsigned long long range_hi(int a, int b) {
    unsigned long long m = 0;
    for (int i = a; i < b; ++i) m |= ohi(i);
    return m;
}

typedef StaticIO<FM_MET6, SAVE_SNOW_DEFAULT(olo) | SAVE_PWATER_DEFAULT(olo), SAVE_SNOW_DEFAULT(ohi) | SAVE_PWATER_DEFAULT(ohi)>
    IO_SnowPwaterDefault;
typedef StaticIO<FM_MET6, SAVE_SNOW_DEFAULT(olo) | SAVE_PWATER_DEFAULT(olo), SAVE_SNOW_DEFAULT(ohi) | SAVE_PWATER_DEFAULT(ohi), true>
    IO_SnowPwaterDefaultF32;
typedef StaticIO<FM_MET6, range_lo(O_SNOW_ALBEDO, O_PWATER_INFFAC + 1), range_hi(O_SNOW_ALBEDO, O_PWATER_INFFAC + 1)>
    IO_SnowPwaterAll;
typedef StaticIO<FM_MET6, SAVE_PWATER_FLOWS(olo), SAVE_PWATER_FLOWS(ohi)> IO_SnowPwaterFlows;
typedef StaticIO<FM_MET6, SAVE_SNOW_DEFAULT(olo) | SAVE_IWATER_DEFAULT(olo), SAVE_SNOW_DEFAULT(ohi) | SAVE_IWATER_DEFAULT(ohi)>
    IO_SnowIwaterDefault;

// full PERLND chain (BASELINE.json configs[3]): every output of the six sections, or the ten series a CBP land-segment
// UCI sends to its river model (tests/land_spec/hwmA51800.uci:493-502)
#define SAVE_CBP10(B)                                                                                             \
    (B(O_PWATER_SURO) | B(O_PWATER_IFWO) | B(O_PWATER_AGWO) | B(O_SEDMNT_SOSED) | B(O_PWTGAS_SOHT) |              \
     B(O_PWTGAS_IOHT) | B(O_PWTGAS_AOHT) | B(O_PWTGAS_SODOXM) | B(O_PWTGAS_IODOXM) | B(O_PWTGAS_AODOXM))
typedef StaticIO<FM_MET6_GATMP, range_lo(O_ATEMP_AIRTMP, O_IWATER_IMPEV), range_hi(O_ATEMP_AIRTMP, O_IWATER_IMPEV)>
    IO_FullChainAll;
typedef StaticIO<FM_MET6_GATMP, SAVE_CBP10(olo), SAVE_CBP10(ohi)> IO_FullChainCbp10;

// does the launch record describe exactly this static configuration (rows in ascending id order)?
template <class IO>
static inline bool io_matches(const LandLaunch &L) {
    if (L.plain || L.forc_f32 || L.linked) return false;  // the library's static kernels are built for the TILED float64 layout
    if (L.nf != IO::NF || L.ns != IO::NS || (L.out_f32 != 0) != IO::F32) return false;
    int r = 0;
    for (int id = 0; id < HSP2G_NFORC; ++id) {
        const bool on = (IO::FMASK >> id) & 1ull;
        if (L.frow[id] != (on ? r : -1)) return false;
        r += on;
    }
    r = 0;
    for (int id = 0; id < HSP2G_NOUT; ++id) {
        const bool on = io_saved<IO>(id);
        if (L.srow[id] != (on ? r : -1)) return false;
        r += on;
    }
    return true;
}

// layout policy of the library's own kernels: static row sets are TILED float64, the general kernels follow the launch record
template <class IO>
struct LayFor {
    typedef TiledLay type;
};
template <>
struct LayFor<DynIO> {
    typedef DynLay type;
};

// option word of test10's PERLND 1 with SNOW + PWATER (ICEFG, CSNOFG, PWATER sees ICEFG; CEPSC, UZSN, NSUR, LZETP follow their
// monthly tables; everything else off: energy-balance snow, Newton routing, UZINF1, IFFCFG 1, VLEFG 1)
#define FLAGW_TEST10_P001 (FLAG(ICEFG) | FLAG(CSNOFG) | FLAG(PW_ICEFG) | FLAG(M_LZETP) | FLAG(M_CEPSC) | FLAG(M_NSUR) | FLAG(M_UZSN))

// ... and of test10's IMPLND 1 with SNOW + IWATER (ICEFG, CSNOFG, RTLIFG; constant RETSC / NSUR, Newton routing)
#define FLAGW_TEST10_I001 (FLAG(ICEFG) | FLAG(CSNOFG) | FLAG(RTLIFG))

// PQUAL / IQUAL rows with the inputs a land batch produces (+ PREC) and every output of the constituent
#define FM_PQUAL7 ((1ull << F_SURO) | (1ull << F_IFWO) | (1ull << F_AGWO) | (1ull << F_PERO) | (1ull << F_WSSD) | \
                   (1ull << F_SCRSD) | (1ull << F_PREC))
#define FM_IQUAL3 ((1ull << F_SURO) | (1ull << F_SOSLD) | (1ull << F_PREC))
constexpr unsigned long long out_range_mask(int lo, int n, int word) {  // ids [lo, lo+n) that fall in 64-bit word `word`
    unsigned long long m = 0;
    for (int id = lo; id < lo + n; ++id)
        if (id / 64 == word) m |= 1ull << (id & 63);
    return m;
}
typedef StaticIO<FM_PQUAL7, 0, out_range_mask(O_PQUAL_SQO, 20, 1), false, out_range_mask(O_PQUAL_SQO, 20, 2)> IO_PqualAll;
typedef StaticIO<FM_IQUAL3, 0, out_range_mask(O_IQUAL_SOQUAL, 12, 1), false, out_range_mask(O_IQUAL_SOQUAL, 12, 2)> IO_IqualAll;
