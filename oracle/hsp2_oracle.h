/*
 * hsp2_oracle.h -- CPU ORACLE for the HSP2 land-segment hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A plain-C restatement (one segment x all timesteps, strictly sequential in time) of the reference's
 * Numba cores in respec/HSPsquared v0.11.0a1:
 *     _atemp_  src/hsp2/hsp2/ATEMP.py:33-59        _snow_   src/hsp2/hsp2/SNOW.py:91-598
 *     _pwater_ src/hsp2/hsp2/PWATER.py:102-769     _iwater_ src/hsp2/hsp2/IWATER.py:75-284
 *     _sedmnt_ src/hsp2/hsp2/SEDMNT.py:52-264      _pstemp_ src/hsp2/hsp2/PSTEMP.py:62-217
 *     _pwtgas_ src/hsp2/hsp2/PWTGAS.py:65-352
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  Nothing under hspsquared_b200/ links, imports or executes it.
 *
 * PARITY PIN: oracle/gen_golden.py runs the UNMODIFIED reference kernels (imported from /root/reference in
 * the build container) and tests/test_oracle_golden.py checks this restatement against those vectors
 * bit-for-bit (same glibc libm, no FMA contraction, same evaluation order), and against the HSPF known-answer
 * files the reference's own tests use (tests/ipwater/data/Flow_Test.plt, tests/test10/HSPFresults/test10I.hbn).
 *
 * Scope notes: English units, and metric (uunits == 2) for SNOW, PWATER and IWATER (SNOW.py:114-160,569-585 here; the array and
 * scalar conversions of PWATER.py:193-244,660-688 / IWATER.py:86-131,276-283 are elementwise and live in oracle/wrappers.py); PWATER HWTFG=1 is unsupported in the reference itself
 * (PWATER.py:107-110) and is refused by the callers.
 *
 * Interface: every routine takes three homogeneous records whose members are listed by the X-macros below:
 *   *_ui  : doubles   (the reference's `ui` Dict[str,float64]; flags are doubles cast with (int))
 *   *_in  : const double* per-step arrays (the reference's `ts` inputs, length >= nsteps)
 *   *_out : double* per-step arrays, length nsteps, fully written
 * hsp2o_fields("snow_ui") etc. returns the comma separated member list so that the ctypes binding in
 * oracle/oracle.py is generated from this header's single source of truth.
 */
#ifndef HSP2_ORACLE_H
#define HSP2_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ATEMP ---------------------------------------------------------------------------------- */
#define ATEMP_UI(X) X(k) X(ELDAT)
#define ATEMP_IN(X) X(LAPSE) X(PREC) X(GATMP)
#define ATEMP_OUT(X) X(AIRTMP)

/* ---- SNOW ----------------------------------------------------------------------------------- */
#define SNOW_UI(X) X(delt) X(cloudfg) X(ICEFG) X(SNOPFG) X(MELEV) X(PACKF) X(PACKI) X(PACKW) X(PAKTMP) \
    X(RDCSN) X(RDENPF) X(SKYCLR) X(TBASE) X(TSNOW) X(COVINX) X(DULL) X(XLNMLT) X(uunits)
#define SNOW_IN(X) X(AIRTMP) X(CCFACT) X(CLOUD) X(COVIND) X(DTMPG) X(HR6IND) X(HRFG) X(KMELT) X(MGMELT) \
    X(MWATER) X(PREC) X(SEASONS) X(SHADE) X(SNOEVP) X(SNOWCF) X(SOLRAD) X(WINMOV)
#define SNOW_OUT(X) X(ALBEDO) X(COVINX) X(DEWTMP) X(DULL) X(MELT) X(NEGHTS) X(PACKF) X(PACKI) X(PACKW) \
    X(PACK) X(PAKTMP) X(PDEPTH) X(PRAIN) X(RAINF) X(RDENPF) X(SKYCLR) X(SNOCOV) X(SNOTMP) X(SNOWE) \
    X(SNOWF) X(WYIELD) X(XLNMLT)

/* ---- PWATER --------------------------------------------------------------------------------- */
#define PWATER_UI(X) X(delt) X(ICEFG) X(CSNOFG) X(IFFCFG) X(IFRDFG) X(RTOPFG) X(UZFG) X(VLEFG) \
    X(AGWETP) X(AGWS) X(BASETP) X(CEPS) X(GWVS) X(IFWS) X(INFEXP) X(INFILD) X(LSUR) X(LZS) X(SLSUR) \
    X(SURS) X(UZS) X(FOREST) X(FZG) X(FZGL) X(INFILT_MFAC)
#define PWATER_IN(X) X(AGWLI) X(AGWRC) X(AIRTMP) X(CEPSC) X(DEEPFR) X(IFWLI) X(INFILT) X(INTFW) X(IRC) \
    X(KVARY) X(LGTMP) X(LZETP) X(LZLI) X(LZSN) X(NSUR) X(PACKI) X(PETINP) X(PETMAX) X(PETMIN) X(PREC) \
    X(RAINF) X(SNOCOV) X(SURLI) X(UZLI) X(UZSN) X(WYIELD) X(DAYFG) X(HRFG)
#define PWATER_OUT(X) X(AGWET) X(AGWI) X(AGWO) X(AGWS) X(BASET) X(CEPE) X(CEPS) X(GWVS) X(IFWI) X(IFWO) \
    X(IFWS) X(IGWI) X(INFIL) X(LZET) X(LZI) X(LZS) X(PERC) X(PERO) X(PERS) X(PET) X(PETADJ) X(SUPY) \
    X(SURI) X(SURO) X(SURS) X(TAET) X(TGWS) X(UZET) X(UZI) X(UZS) X(INFFAC)

/* ---- IWATER --------------------------------------------------------------------------------- */
#define IWATER_UI(X) X(delt) X(CSNOFG) X(RTOPFG) X(RTLIFG) X(LSUR) X(SLSUR) X(RETS) X(SURS)
#define IWATER_IN(X) X(HRFG) X(HR1FG) X(RETSC) X(NSUR) X(AIRTMP) X(PETMAX) X(PETMIN) X(SNOCOV) X(SURLI) \
    X(PETINP) X(RAINF) X(WYIELD) X(PREC)
#define IWATER_OUT(X) X(IMPEV) X(PET) X(PETADJ) X(RETS) X(SUPY) X(SURI) X(SURO) X(SURS)

/* ---- SEDMNT --------------------------------------------------------------------------------- */
#define SEDMNT_UI(X) X(delt) X(VSIVFG) X(SDOPFG) X(CSNOFG) X(SMPF) X(KRER) X(JRER) X(AFFIX) X(NVSI) \
    X(KSER) X(JSER) X(KGER) X(JGER) X(DETS) X(uunits)
#define SEDMNT_IN(X) X(COVERI) X(NVSI) X(RAIN) X(PREC) X(SLSED) X(SNOCOV) X(SURO) X(SURS) X(DAYFG)
#define SEDMNT_OUT(X) X(DETS) X(WSSD) X(SCRSD) X(SOSED) X(COVER)

/* ---- PSTEMP --------------------------------------------------------------------------------- */
#define PSTEMP_UI(X) X(TSOPFG) X(AIRTFG) X(AIRTC) X(SLTMP) X(ULTMP) X(LGTMP) X(uunits)
#define PSTEMP_IN(X) X(AIRTMP) X(ASLT) X(BSLT) X(ULTP1) X(ULTP2) X(LGTP1) X(LGTP2) X(HRFG)
#define PSTEMP_OUT(X) X(AIRTC) X(SLTMP) X(ULTMP) X(LGTMP)

/* ---- PWTGAS --------------------------------------------------------------------------------- */
#define PWTGAS_UI(X) X(CSNOFG) X(SLIFAC) X(ILIFAC) X(ALIFAC) X(ELEV) X(GDVFG) X(uunits)
#define PWTGAS_IN(X) X(IDOXP) X(ICO2P) X(ADOXP) X(ACO2P) X(WYIELD) X(SURO) X(IFWO) X(AGWO) X(SURLI) \
    X(IFWLI) X(AGWLI) X(SLTMP) X(ULTMP) X(LGTMP) X(SLITMP) X(SLIDOX) X(SLICO2) X(ILITMP) X(ILIDOX) \
    X(ILICO2) X(ALITMP) X(ALIDOX) X(ALICO2) X(DAYFG)
#define PWTGAS_OUT(X) X(SOTMP) X(IOTMP) X(AOTMP) X(SODOX) X(SOCO2) X(IODOX) X(IOCO2) X(AODOX) X(AOCO2) \
    X(SOHT) X(IOHT) X(AOHT) X(POHT) X(SODOXM) X(SOCO2M) X(IODOXM) X(IOCO2M) X(AODOXM) X(AOCO2M) \
    X(PODOXM) X(POCO2M)

/* ---- SOLIDS (IMPLND) ------------------------------------------------------------------------- */
#define SOLIDS_UI(X) X(delt60) X(SDOPFG) X(KEIM) X(JEIM) X(SLDS)
#define SOLIDS_IN(X) X(SURO) X(SURS) X(PREC) X(ACCSDP) X(REMSDP) X(DAYFG)
#define SOLIDS_OUT(X) X(SOSLD) X(SLDS)

/* ---- IWTGAS (IMPLND) ------------------------------------------------------------------------- */
#define IWTGAS_UI(X) X(SLIFAC) X(SOTMP) X(SODOX) X(SOCO2) X(ELEV) X(uunits)
#define IWTGAS_IN(X) X(AIRTMP) X(WYIELD) X(SURO) X(SURLI) X(SLITMP) X(SLIDOX) X(SLICO2) X(AWTF) X(BWTF) X(DAYFG)
#define IWTGAS_OUT(X) X(SOTMP) X(SODOX) X(SOCO2) X(SOHT) X(SODOXM) X(SOCO2M)

/* ---- PQUAL (PERLND) / IQUAL (IMPLND): ONE constituent per call; the reference loops constituents inside its kernel
 *      with per-constituent ui keys (PQUAL.py:148-170, IQUAL.py:131-150) and oracle.py does that loop ------------- */
#define PQUAL_UI(X) X(delt60) X(QSDFG) X(QSOFG) X(QIFWFG) X(QAGWFG) X(VIQCFG) X(VAQCFG) X(SQO) X(WSQOP) X(SLIFAC) \
    X(ILIFAC) X(ALIFAC)
#define PQUAL_IN(X) X(SURO) X(IFWO) X(AGWO) X(PERO) X(WSSD) X(SCRSD) X(PREC) X(SLIQSP) X(ILIQC) X(ALIQC) X(POTFW) \
    X(POTFS) X(ACQOP) X(REMQOP) X(IOQCP) X(AOQCP) X(PQADFX) X(PQADCN) X(DAYFG)
#define PQUAL_OUT(X) X(SQO) X(SOQSP) X(SOQOC) X(SOQC) X(IOQC) X(AOQC) X(POQC) X(WASHQS) X(SCRQS) X(SOQS) X(SOQO) \
    X(SOQUAL) X(IOQUAL) X(AOQUAL) X(POQUAL) X(PQADDR) X(PQADWT) X(PQADEP) X(SLIQO) X(INFLOW)
#define IQUAL_UI(X) X(delt60) X(QSDFG) X(QSOFG) X(SQO) X(WSQOP) X(SLIFAC)
#define IQUAL_IN(X) X(SURO) X(SOSLD) X(PREC) X(SLIQSP) X(POTFW) X(ACQOP) X(REMQOP) X(IQADFX) X(IQADCN) X(DAYFG)
#define IQUAL_OUT(X) X(SOQUAL) X(SOQC) X(SOQO) X(SQO) X(SOQOC) X(SOQS) X(SOQSP) X(IQADDR) X(IQADWT) X(IQADEP) \
    X(SLIQO) X(INFLOW)

#define HSP2O_DECL_UI(n) double n;
#define HSP2O_DECL_IN(n) const double *n;
#define HSP2O_DECL_OUT(n) double *n;

#define HSP2O_RECORDS(NAME, UI, IN, OUT)             \
    typedef struct { UI(HSP2O_DECL_UI) } NAME##_ui;   \
    typedef struct { IN(HSP2O_DECL_IN) } NAME##_in;   \
    typedef struct { OUT(HSP2O_DECL_OUT) } NAME##_out;

HSP2O_RECORDS(atemp, ATEMP_UI, ATEMP_IN, ATEMP_OUT)
HSP2O_RECORDS(snow, SNOW_UI, SNOW_IN, SNOW_OUT)
HSP2O_RECORDS(pwater, PWATER_UI, PWATER_IN, PWATER_OUT)
HSP2O_RECORDS(iwater, IWATER_UI, IWATER_IN, IWATER_OUT)
HSP2O_RECORDS(sedmnt, SEDMNT_UI, SEDMNT_IN, SEDMNT_OUT)
HSP2O_RECORDS(pstemp, PSTEMP_UI, PSTEMP_IN, PSTEMP_OUT)
HSP2O_RECORDS(pwtgas, PWTGAS_UI, PWTGAS_IN, PWTGAS_OUT)
HSP2O_RECORDS(solids, SOLIDS_UI, SOLIDS_IN, SOLIDS_OUT)
HSP2O_RECORDS(iwtgas, IWTGAS_UI, IWTGAS_IN, IWTGAS_OUT)
HSP2O_RECORDS(pqual, PQUAL_UI, PQUAL_IN, PQUAL_OUT)
HSP2O_RECORDS(iqual, IQUAL_UI, IQUAL_IN, IQUAL_OUT)

/* error-count vector lengths (ERRMSGS tuples: SNOW.py:23, PWATER.py:15-25, IWATER.py:15, PSTEMP.py:13-18) */
#define HSP2O_NERR_SNOW 1
#define HSP2O_NERR_PWATER 10
#define HSP2O_NERR_IWATER 1
#define HSP2O_NERR_PSTEMP 6

const char *hsp2o_fields(const char *record_name);

void hsp2o_atemp(const atemp_ui *ui, const atemp_in *in, const atemp_out *out, int64_t nsteps);
void hsp2o_snow(const snow_ui *ui, const snow_in *in, const snow_out *out, int64_t nsteps, int64_t *errors);
void hsp2o_pwater(const pwater_ui *ui, const pwater_in *in, const pwater_out *out, int64_t nsteps, int64_t *errors);
void hsp2o_iwater(const iwater_ui *ui, const iwater_in *in, const iwater_out *out, int64_t nsteps, int64_t *errors);
void hsp2o_sedmnt(const sedmnt_ui *ui, const sedmnt_in *in, const sedmnt_out *out, int64_t nsteps);
void hsp2o_pstemp(const pstemp_ui *ui, const pstemp_in *in, const pstemp_out *out, int64_t nsteps, int64_t *errors);
void hsp2o_pwtgas(const pwtgas_ui *ui, const pwtgas_in *in, const pwtgas_out *out, int64_t nsteps);
void hsp2o_solids(const solids_ui *ui, const solids_in *in, const solids_out *out, int64_t nsteps);
void hsp2o_iwtgas(const iwtgas_ui *ui, const iwtgas_in *in, const iwtgas_out *out, int64_t nsteps);
void hsp2o_pqual(const pqual_ui *ui, const pqual_in *in, const pqual_out *out, int64_t nsteps);
void hsp2o_iqual(const iqual_ui *ui, const iqual_in *in, const iqual_out *out, int64_t nsteps);

/*
 * Multi-threaded CPU baseline over independent segments (bench.py cpu_baseline / --impl reference).
 * Segment-major layout, i.e. the reference's natural one: every per-step array of segment s is contiguous.
 *   forc  [nseg][6][nsteps]   PREC, PETINP, AIRTMP, WINMOV, SOLRAD, DTMPG
 *   par   [nseg][HSP2O_BATCH_NPAR] scalar parameters / initial states (order: hsp2o_fields("batch_par"))
 *   mon   [nseg][4][nsteps+1] per-step CEPSC, UZSN, NSUR, LZETP series (already daily-interpolated)
 *   cal   [4][nsteps+1]       HRFG, DAYFG, HR6IND, SEASONS (shared)
 *   sums  [nseg][24]          time-sum of each default-SAVE series (a checksum so the work cannot be elided)
 * Each thread keeps its own set of 22+31 output arrays of length nsteps and rewrites them per segment,
 * which is what the reference does when it allocates ts[...] per segment.
 * Returns the number of threads used.
 */
#define BATCH_PAR(X) X(MELEV) X(SHADE) X(SNOWCF) X(COVIND) X(KMELT) X(TBASE) X(RDCSN) X(TSNOW) X(SNOEVP) \
    X(CCFACT) X(MWATER) X(MGMELT) X(PACKF) X(PACKI) X(PACKW) X(RDENPF) X(DULL) X(PAKTMP) X(COVINX) \
    X(XLNMLT) X(SKYCLR) X(ICEFG) X(SNOPFG) X(FOREST) X(LZSN) X(INFILT) X(LSUR) X(SLSUR) X(KVARY) \
    X(AGWRC) X(PETMAX) X(PETMIN) X(INFEXP) X(INFILD) X(DEEPFR) X(BASETP) X(AGWETP) X(INTFW) X(IRC) \
    X(FZG) X(FZGL) X(CEPS) X(SURS) X(UZS) X(IFWS) X(LZS) X(AGWS) X(GWVS) X(RTOPFG) X(UZFG) X(VLEFG) \
    X(IFFCFG) X(IFRDFG)
enum {
#define X(n) BP_##n,
    BATCH_PAR(X)
#undef X
        HSP2O_BATCH_NPAR
};
int hsp2o_perlnd_batch(int64_t nseg, int64_t nsteps, double delt, const double *forc, const double *par,
                       const double *mon, const double *cal, double *sums, int64_t *errors /*[nseg][11]*/,
                       int nthreads);

#ifdef __cplusplus
}
#endif
#endif
