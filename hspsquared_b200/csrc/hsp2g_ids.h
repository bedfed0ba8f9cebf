/*
 * hsp2g_ids.h -- identifier spaces of the land-segment engine (shared by the CUDA kernels and the C ABI).
 *
 * Every list is an X-macro so that the library can export its own name tables (hsp2g_names()), from which the
 * Python host code builds its index maps: one source of truth, no hand-mirrored enums.
 *
 * Names follow the HSPF/HSP2 domain: the parameter / state / flux names are the reference's `ui` / `ts` keys
 * (src/hsp2/hsp2/{ATEMP,SNOW,PWATER,IWATER,SEDMNT,PSTEMP,PWTGAS}.py; tables in hsp2tools/data/ParseTable.csv).
 */
#ifndef HSP2G_IDS_H
#define HSP2G_IDS_H

/* ---- operations and activities ------------------------------------------------------------------ */
#define HSP2G_OP_PERLND 0
#define HSP2G_OP_IMPLND 1

/* activity mask bits, registry order of configuration.py:48-52 */
#define HSP2G_ACT_ATEMP 1u
#define HSP2G_ACT_SNOW 2u
#define HSP2G_ACT_PWATER 4u
#define HSP2G_ACT_SEDMNT 8u
#define HSP2G_ACT_PSTEMP 16u
#define HSP2G_ACT_PWTGAS 32u
#define HSP2G_ACT_IWATER 64u
#define HSP2G_ACT_SOLIDS 128u /* IMPLND, after IWATER (configuration.py:51-52) */
#define HSP2G_ACT_IWTGAS 256u
/* quality constituents: a batch with ONLY this bit set holds one row per (segment, constituent) and reads the flows and
 * sediment / solids outflows it is associated with as input series (PQUAL.py:147, IQUAL.py:109) */
#define HSP2G_ACT_PQUAL 512u
#define HSP2G_ACT_IQUAL 1024u

/* ---- per-segment scalar parameters: par[P_x][segment], fp64 -------------------------------------- */
/* derived columns (computed by the host with the reference's own expression order):
 *   MGMELT_IVL = MGMELT*delt/1440 (SNOW.py:146)   INFILT_IVL = INFILT*delt60 (PWATER.py:215)
 *   KGW = 1 - AGWRC**(delt60/24) (PWATER.py:247)   ELEVGC = ((288-0.00198*ELEV)/288)**5.256 (PWTGAS.py:95)
 *   the per-segment constant heads of three SNOW expressions, each the FIRST operation(s) of the reference's left-to-right
 *   evaluation, so that hoisting them changes no bit:  MELEV_FAC = 1.0 - 0.3*MELEV/10000.0 (SNOW.py:360, one fp64 division per
 *   interval with a pack otherwise),  SNOEVP_K = SNOEVP*0.0002 (SNOW.py:336),  CCFACT_K = CCFACT*0.00026 (SNOW.py:354) */
#define HSP2G_PARAMS(X)                                                                                   \
    X(ELDAT)                                                                                              \
    X(MELEV_FAC) X(SHADE) X(SNOWCF) X(COVIND) X(KMELT) X(TBASE) X(RDCSN) X(TSNOW) X(SNOEVP_K) X(CCFACT_K) \
    X(MWATER) X(MGMELT_IVL)                                                                               \
    X(FOREST) X(LZSN) X(INFILT_IVL) X(LSUR) X(SLSUR) X(KVARY) X(KGW) X(PETMAX) X(PETMIN) X(INFEXP)        \
    X(INFILD) X(DEEPFR) X(BASETP) X(AGWETP) X(CEPSC) X(UZSN) X(NSUR) X(INTFW) X(IRC) X(LZETP) X(FZG)      \
    X(FZGL)                                                                                               \
    X(SMPF) X(KRER) X(JRER) X(AFFIX) X(COVER) X(NVSI) X(KSER) X(JSER) X(KGER) X(JGER)                     \
    X(ASLT) X(BSLT) X(ULTP1) X(ULTP2) X(LGTP1) X(LGTP2)                                                   \
    X(ELEVGC) X(IDOXP) X(ICO2P) X(ADOXP) X(ACO2P) X(SLIFAC) X(ILIFAC) X(ALIFAC)                           \
    X(RETSC)                                                                                              \
    /* IMPLND SOLIDS (SOLIDS.py:62-64) and IWTGAS (IWTGAS.py:76-86; ELEVGC and SLIFAC share the PWTGAS columns) */ \
    X(KEIM) X(JEIM) X(ACCSDP) X(REMSDP) X(AWTF) X(BWTF)                                                   \
    /* PQUAL / IQUAL rows: Q_WSFAC = 2.30 / WSQOP (PQUAL.py:166, IQUAL.py:147); Q_REMQOP = ACQOP / SQOLIM (initmdiv,  \
     * utilities.py:214-225); Q_ADFX / Q_ADCN = atmospheric deposition flux / concentration (0 unless monthly or a series) */ \
    X(Q_WSFAC) X(Q_POTFW) X(Q_POTFS) X(Q_ACQOP) X(Q_REMQOP) X(Q_IOQC) X(Q_AOQC) X(Q_ADFX) X(Q_ADCN)

/* ---- monthly tables: mon[slot][12][segment]; a parameter follows its table only where the segment's
 *      M_<name> flag bit is set (utilities.py:205-211 initm: flag truthy AND table present) -------------- */
#define HSP2G_MONTHLY(X)                                                                                  \
    X(KMELT) X(LZETP) X(CEPSC) X(INTFW) X(IRC) X(NSUR) X(UZSN) X(COVER) X(NVSI) X(ASLT) X(BSLT) X(ULTP1)  \
    X(ULTP2) X(LGTP1) X(LGTP2) X(IDOXP) X(ICO2P) X(ADOXP) X(ACO2P) X(RETSC) X(ACCSDP) X(REMSDP) X(AWTF) X(BWTF) \
    X(Q_POTFW) X(Q_POTFS) X(Q_ACQOP) X(Q_REMQOP) X(Q_IOQC) X(Q_AOQC) X(Q_ADFX) X(Q_ADCN)

/* ---- per-segment option flags: one 64-bit word per segment ---------------------------------------- */
#define HSP2G_FLAGS(X)                                                                                    \
    X(ICEFG)      /* SNOW ICEFG (SNOW.py:105-107) */                                                      \
    X(SNOPFG)     /* degree-day snowmelt (SNOW.py:125-127) */                                             \
    X(CSNOFG)     /* PWATER / IWATER consider snow (PWATER.py:170, IWATER.py:91) */                        \
    X(PW_ICEFG)   /* ICEFG as PWATER sees it through siminfo (SNOW.py:60-63, PWATER.py:87) */             \
    X(RTOP1)      /* RTOPFG == 1: ARM/NPS routing (PWATER.py:701); IWATER: RTOPFG truthy (IWATER.py:202) */ \
    X(UZFG) X(IFRDFG)                                                                                     \
    X(IFFCFG2)    /* IFFCFG == 2 (PWATER.py:302) */                                                       \
    X(VLE_GE2)    /* VLEFG >= 2 (PWATER.py:615-620) */                                                    \
    X(RTLIFG)     /* IWATER lateral inflow subject to retention (IWATER.py:174) */                         \
    X(SED_CSNO1)  /* SEDMNT: CSNOFG == 1 (SEDMNT.py:164) */                                               \
    X(SDOPFG) X(VSIV_EQ2) X(VSIV_LE2)                                                                     \
    X(TSOP1) X(TSOP0) X(AIRTFG)                                                                           \
    X(PWG_CSNO)   /* PWTGAS: CSNOFG truthy (PWTGAS.py:203) */                                              \
    X(GDVFG)                                                                                              \
    X(M_KMELT) X(M_LZETP) X(M_CEPSC) X(M_INTFW) X(M_IRC) X(M_NSUR) X(M_UZSN) X(M_COVER) X(M_NVSI)         \
    X(M_ASLT) X(M_BSLT) X(M_ULTP1) X(M_ULTP2) X(M_LGTP1) X(M_LGTP2) X(M_IDOXP) X(M_ICO2P) X(M_ADOXP)      \
    X(M_ACO2P) X(M_RETSC)                                                                                 \
    X(SLD_SDOP1)  /* SOLIDS: SDOPFG == 1 (SOLIDS.py:103) */                                               \
    X(M_ACCSDP) X(M_REMSDP) X(M_AWTF) X(M_BWTF)                                                           \
    /* PQUAL / IQUAL rows: QSDFG; QSOFG != 0, >= 1, == 1, == 2, == -1; QIFWFG; QAGWFG; VIQCFG / VAQCFG in (3, 4) */    \
    X(Q_QSDFG) X(Q_QSO_NZ) X(Q_QSO_GE1) X(Q_QSO_EQ1) X(Q_QSO_EQ2) X(Q_QSO_M1) X(Q_QIFWFG) X(Q_QAGWFG) X(Q_VIQC34)  \
    X(Q_VAQC34)                                                                                           \
    X(M_Q_POTFW) X(M_Q_POTFS) X(M_Q_ACQOP) X(M_Q_REMQOP) X(M_Q_IOQC) X(M_Q_AOQC) X(M_Q_ADFX) X(M_Q_ADCN)

/* ---- carried state: state[S_x][segment], fp64.  Named states (restart.py:9-14) AND the hidden scalars the
 *      reference carries across steps, so a run split into time chunks is bit-identical
 *      to a single pass.  Integer latches are stored as 0.0/1.0. ------------------------------------------ */
#define HSP2G_STATES(X)                                                                                   \
    /* SNOW */                                                                                            \
    X(COVINX) X(DULL) X(PACKF) X(PACKI) X(PACKW) X(PAKTMP) X(RDENPF) X(SKYCLR) X(XLNMLT) X(PDEPTH)        \
    X(SNOCOV) X(NEGHTS) X(OLDPREC) X(SNOTMP) X(DEWTMP) X(MNEGHS) X(PACKWC) X(COMPCT) X(GMELTR) X(MOSTHT)  \
    X(NEGHT) X(RDNSN) X(SNOWEP) X(VAP) X(ALBEDO) X(HR6FG)                                                 \
    /* PWATER */                                                                                          \
    X(CEPS) X(SURS) X(UZS) X(IFWS) X(LZS) X(AGWS) X(GWVS) X(MSUPY) X(DEC) X(SRC) X(IFWK1) X(IFWK2)        \
    X(RLZRAT) X(LZFRAC) X(RPARM) X(PETADJ)                                                                \
    /* SEDMNT */                                                                                          \
    X(DETS) X(SED_COVER) X(SED_NVSI) X(DRYDFG)                                                            \
    /* PSTEMP (deg C internally) */                                                                       \
    X(SLTMP) X(ULTMP) X(LGTMP) X(AIRTS)                                                                   \
    /* IWATER (SURS, MSUPY, DEC, SRC, PETADJ shared with the PWATER slots: a batch is one operation) */    \
    X(RETS)                                                                                               \
    /* SOLIDS: storage and the dry-day latch; IWTGAS: the day's temperature regression coefficients */    \
    X(SLDS) X(SLD_DRYDFG) X(IWG_AWTF) X(IWG_BWTF)                                                         \
    /* PQUAL / IQUAL: surface storage; IQUAL's carried washoff potency (IQUAL.py:188, 207-214) */          \
    X(Q_SQO) X(Q_SOQSP)

/* ---- forcing / input series ids: forc[t][row][segment]; absent ids read a per-batch fill value that the
 *      host derives by replaying the reference wrappers' NaN / zero / -1e30 fills (SNOW.py:44-46,
 *      PWATER.py:41-48, IWATER.py:35-42, SEDMNT.py:94-96, PWTGAS.py:106-140) ---------------------------- */
#define HSP2G_FORCINGS(X)                                                                                 \
    X(GATMP) X(PREC) X(PETINP) X(AIRTMP) X(WINMOV) X(SOLRAD) X(DTMPG) X(CLOUD)                            \
    X(SURLI) X(UZLI) X(IFWLI) X(LZLI) X(AGWLI) X(LGTMP)                                                   \
    X(RAINF) X(SNOCOV) X(WYIELD) X(PACKI)                                                                 \
    X(SLSED) X(SURO) X(SURS) X(IFWO) X(AGWO) X(SLTMP) X(ULTMP)                                            \
    X(SLITMP) X(SLIDOX) X(SLICO2) X(ILITMP) X(ILIDOX) X(ILICO2) X(ALITMP) X(ALIDOX) X(ALICO2)             \
    /* inputs of PQUAL / IQUAL rows (PQUAL.py:119-126,208-210, IQUAL.py:118-124); QADFX / QADCN = the constituent's       \
     * deposition series when its PQADFG / IQADFG flag is -1 */                                                          \
    X(PERO) X(WSSD) X(SCRSD) X(SOSLD) X(SLIQSP) X(ILIQC) X(ALIQC) X(QADFX) X(QADCN)

/* ---- output series ids (the reference's ts[...] outputs; SAVE selects the rows that are written) ------ */
#define HSP2G_OUT_ATEMP(X) X(AIRTMP)
#define HSP2G_OUT_SNOW(X)                                                                                 \
    X(ALBEDO) X(COVINX) X(DEWTMP) X(DULL) X(MELT) X(NEGHTS) X(PACKF) X(PACKI) X(PACKW) X(PACK) X(PAKTMP)  \
    X(PDEPTH) X(PRAIN) X(RAINF) X(RDENPF) X(SKYCLR) X(SNOCOV) X(SNOTMP) X(SNOWE) X(SNOWF) X(WYIELD)       \
    X(XLNMLT)
#define HSP2G_OUT_PWATER(X)                                                                               \
    X(AGWET) X(AGWI) X(AGWO) X(AGWS) X(BASET) X(CEPE) X(CEPS) X(GWVS) X(IFWI) X(IFWO) X(IFWS) X(IGWI)     \
    X(INFIL) X(LZET) X(LZI) X(LZS) X(PERC) X(PERO) X(PERS) X(PET) X(PETADJ) X(SUPY) X(SURI) X(SURO)       \
    X(SURS) X(TAET) X(TGWS) X(UZET) X(UZI) X(UZS) X(INFFAC)
#define HSP2G_OUT_SEDMNT(X) X(DETS) X(WSSD) X(SCRSD) X(SOSED) X(COVER)
#define HSP2G_OUT_PSTEMP(X) X(AIRTC) X(SLTMP) X(ULTMP) X(LGTMP)
#define HSP2G_OUT_PWTGAS(X)                                                                               \
    X(SOTMP) X(IOTMP) X(AOTMP) X(SODOX) X(SOCO2) X(IODOX) X(IOCO2) X(AODOX) X(AOCO2) X(SOHT) X(IOHT)      \
    X(AOHT) X(POHT) X(SODOXM) X(SOCO2M) X(IODOXM) X(IOCO2M) X(AODOXM) X(AOCO2M) X(PODOXM) X(POCO2M)
/* IWATER outputs: IMPEV PET PETADJ RETS SUPY SURI SURO SURS (IWATER.py:119-126) */
#define HSP2G_OUT_IWATER(X) X(IMPEV) X(PET) X(PETADJ) X(RETS) X(SUPY) X(SURI) X(SURO) X(SURS)
#define HSP2G_OUT_SOLIDS(X) X(SOSLD) X(SLDS)                         /* SOLIDS.py:72-73 */
#define HSP2G_OUT_IWTGAS(X) X(SOTMP) X(SODOX) X(SOCO2) X(SOHT) X(SODOXM) X(SOCO2M) /* IWTGAS.py:104-109 */
/* per constituent: ts['PQUAL<k>_<NAME>'] (PQUAL.py:176-206), ts['IQUAL<k>_<NAME>'] (IQUAL.py:152-175) */
#define HSP2G_OUT_PQUAL(X)                                                                                \
    X(SQO) X(SOQSP) X(SOQOC) X(SOQC) X(IOQC) X(AOQC) X(POQC) X(WASHQS) X(SCRQS) X(SOQS) X(SOQO) X(SOQUAL)  \
    X(IOQUAL) X(AOQUAL) X(POQUAL) X(PQADDR) X(PQADWT) X(PQADEP) X(SLIQO) X(INFLOW)
#define HSP2G_OUT_IQUAL(X)                                                                                \
    X(SOQUAL) X(SOQC) X(SOQO) X(SQO) X(SOQOC) X(SOQS) X(SOQSP) X(IQADDR) X(IQADWT) X(IQADEP) X(SLIQO) X(INFLOW)

/* ---- error counters: err[E_x][segment], int64 (ERRMSGS tuples of the reference) ---------------------- */
#define HSP2G_ERRORS(X)                                                                                   \
    X(SNOW0)                                                                                              \
    X(PWATER0) X(PWATER1) X(PWATER2) X(PWATER3) X(PWATER4) X(PWATER5) X(PWATER6) X(PWATER7) X(PWATER8)    \
    X(PWATER9)                                                                                            \
    X(PSTEMP0) X(PSTEMP1) X(PSTEMP2) X(PSTEMP3) X(PSTEMP4) X(PSTEMP5)                                     \
    X(IWATER0)                                                                                            \
    /* PQUAL / IQUAL: deposition on a constituent without QSOFG, counted once per run by the host (PQUAL.py:171-174) */ \
    X(PQUAL0) X(PQUAL1) X(IQUAL0) X(IQUAL1)

/* per-step shared calendar flag bits (hspsquared_b200/calendar.py) */
#define HSP2G_CAL_HRFG 1u
#define HSP2G_CAL_DAYFG 2u
#define HSP2G_CAL_HR1FG 4u
#define HSP2G_CAL_HR6IND 8u
#define HSP2G_CAL_SEASON 16u
#define HSP2G_CAL_NEWDAY 32u

enum {
#define X(n) P_##n,
    HSP2G_PARAMS(X)
#undef X
        HSP2G_NPAR
};
enum {
#define X(n) M_##n,
    HSP2G_MONTHLY(X)
#undef X
        HSP2G_NMON
};
enum {
#define X(n) FB_##n,
    HSP2G_FLAGS(X)
#undef X
        HSP2G_NFLAG
};
enum {
#define X(n) S_##n,
    HSP2G_STATES(X)
#undef X
        HSP2G_NSTATE
};
enum {
#define X(n) F_##n,
    HSP2G_FORCINGS(X)
#undef X
        HSP2G_NFORC
};
enum {
#define X(n) O_ATEMP_##n,
    HSP2G_OUT_ATEMP(X)
#undef X
#define X(n) O_SNOW_##n,
        HSP2G_OUT_SNOW(X)
#undef X
#define X(n) O_PWATER_##n,
            HSP2G_OUT_PWATER(X)
#undef X
#define X(n) O_SEDMNT_##n,
                HSP2G_OUT_SEDMNT(X)
#undef X
#define X(n) O_PSTEMP_##n,
                    HSP2G_OUT_PSTEMP(X)
#undef X
#define X(n) O_PWTGAS_##n,
                        HSP2G_OUT_PWTGAS(X)
#undef X
#define X(n) O_IWATER_##n,
                            HSP2G_OUT_IWATER(X)
#undef X
#define X(n) O_SOLIDS_##n,
                                HSP2G_OUT_SOLIDS(X)
#undef X
#define X(n) O_IWTGAS_##n,
                                    HSP2G_OUT_IWTGAS(X)
#undef X
#define X(n) O_PQUAL_##n,
                                        HSP2G_OUT_PQUAL(X)
#undef X
#define X(n) O_IQUAL_##n,
                                            HSP2G_OUT_IQUAL(X)
#undef X
                                                HSP2G_NOUT
};
enum {
#define X(n) E_##n,
    HSP2G_ERRORS(X)
#undef X
        HSP2G_NERR
};

#define FLAG(n) (1ull << FB_##n)

#endif
