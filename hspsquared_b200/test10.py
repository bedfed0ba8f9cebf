"""The test10 land segments as HSP2 `uci[(operation, activity, segment)]` tables.

Values are transcribed from the HSPF test10 input deck shipped with the reference
(tests/test10/HSP2results/test10.uci:42-169 PERLND 1, :171-248 IMPLND 1) with the defaults / renames /
fix-ups its importer applies (hsp2tools/readUCI.py:209-289, hsp2tools/data/ParseTable.csv):
PKSNOW/PKICE/PKWATR -> PACKF/PACKI/PACKW, SNOPFG=0 when SNOW-FLAGS is absent, FZG=1 / FZGL=0.1 when
PWAT-PARM5 is absent, blank fields -> table defaults.  They are the base of BASELINE.json's synthetic
workloads and of the golden vectors under tests/golden/.  All numbers are native Python
int/float because the reference silently drops anything else (utilities.py:75-78).
"""
import copy

MONTHS = ("JAN", "FEB", "MAR", "APR", "MAY", "JUN", "JUL", "AUG", "SEP", "OCT", "NOV", "DEC")


def monthly(values):
    return dict(zip(MONTHS, (float(v) for v in values)))


_SNOW_STATES = {"PACKF": 1.4, "PACKI": 0.2, "PACKW": 0.1, "RDENPF": 0.2, "DULL": 375.0, "PAKTMP": 27.5,
                "COVINX": 0.50, "XLNMLT": 0.0, "SKYCLR": 1.0}


def _snow(melev):
    return {
        "FLAGS": {"ICEFG": 1, "SNOPFG": 0, "VKMFG": 0},
        "PARAMETERS": {"LAT": 42.0, "MELEV": melev, "SHADE": 0.0, "SNOWCF": 1.45, "COVIND": 0.5, "KMELT": 0.0,
                       "TBASE": 32.0, "RDCSN": 0.12, "TSNOW": 32.0, "SNOEVP": 0.05, "CCFACT": 0.5,
                       "MWATER": 0.08, "MGMELT": 0.0001},
        "STATES": dict(_SNOW_STATES),
    }


# PERLND 1 "BICKNELL FARM": ACTIVITY SNOW, PWATER, PSTEMP, PWTGAS (test10.uci:44-48)
P001 = {
    "ACTIVITY": {"ATEMP": 0, "SNOW": 1, "PWATER": 1, "SEDMNT": 0, "PSTEMP": 1, "PWTGAS": 1},
    "SNOW": _snow(520.0),
    "PWATER": {
        "PARAMETERS": {
            "CSNOFG": 1, "RTOPFG": 0, "UZFG": 0, "VCSFG": 1, "VUZFG": 1, "VNNFG": 1, "VIFWFG": 0, "VIRCFG": 0,
            "VLEFG": 1, "IFFCFG": 1, "HWTFG": 0, "IRRGFG": 0, "IFRDFG": 0,
            "FOREST": 0.010, "LZSN": 8.0, "INFILT": 0.150, "LSUR": 250.0, "SLSUR": 0.050, "KVARY": 0.5,
            "AGWRC": 0.98, "PETMAX": 40.0, "PETMIN": 35.0, "INFEXP": 2.0, "INFILD": 2.0, "DEEPFR": 0.10,
            "BASETP": 0.0, "AGWETP": 0.08, "CEPSC": 0.0, "UZSN": 0.01, "NSUR": 0.1, "INTFW": 1.0, "IRC": 0.60,
            "LZETP": 0.0, "FZG": 1.0, "FZGL": 0.1},
        "STATES": {"CEPS": 0.05, "SURS": 0.0, "UZS": 0.15, "IFWS": 0.0, "LZS": 4.0, "AGWS": 0.05, "GWVS": 0.05},
        "MONTHLY_CEPSC": monthly([0.04, 0.04, 0.03, 0.03, 0.03, 0.03, 0.10, 0.17, 0.19, 0.14, 0.05, 0.04]),
        "MONTHLY_UZSN": monthly([0.4, 0.4, 0.4, 0.4, 1.6, 1.1, 1.1, 1.3, 1.3, 1.3, 1.1, 0.9]),
        "MONTHLY_NSUR": monthly([0.30, 0.30, 0.30, 0.30, 0.27, 0.25, 0.25, 0.25, 0.25, 0.25, 0.35, 0.33]),
        "MONTHLY_LZETP": monthly([0.20, 0.20, 0.20, 0.23, 0.23, 0.25, 0.60, 0.80, 0.75, 0.50, 0.30, 0.20]),
    },
    "PSTEMP": {
        # PSTEMP-PARM1 and PSTEMP-TEMPS are absent from the deck: no flag keys, kernel defaults of 60 F
        "PARAMETERS": {"ASLT": 14.5, "BSLT": 0.365, "ULTP1": 1.2, "ULTP2": 4.0, "LGTP1": 1.2, "LGTP2": 6.0},
    },
    "PWTGAS": {
        # PWT-PARM1 is absent from the deck: no flag keys
        "PARAMETERS": {"ELEV": 500.0, "IDOXP": 6.0, "ICO2P": 0.05,
                       "ADOXP": 5.0, "ACO2P": 0.05, "SDLFAC": 0.0, "SLIFAC": 0.0, "ILIFAC": 0.0, "ALIFAC": 0.0,
                       "SOTMP": 60.0, "IOTMP": 60.0, "AOTMP": 60.0, "SODOX": 0.0, "SOCO2": 0.0, "IODOX": 0.0,
                       "IOCO2": 0.0, "AODOX": 0.0, "AOCO2": 0.0},
    },
}

# IMPLND 1 "DONIGIAN INDUSTRY": ACTIVITY SNOW, IWATER, SOLIDS, IWTGAS (+ IQUAL, out of scope) (test10.uci:173-177).
# IWTGAS is OFF here although the UCI switches it on: the UCI has no IWT-INIT table, readUCI adds no default for it
# (readUCI.py:244-250 only patches LAT-FACTOR), and the reference then raises KeyError('SOTMP') at IWTGAS.py:79 (checked
# against the reference in the build container).  The `i001_iwtgas` parity case supplies the table.
I001 = {
    "ACTIVITY": {"ATEMP": 0, "SNOW": 1, "IWATER": 1, "SOLIDS": 1, "IWTGAS": 0},
    "SNOW": _snow(450.0),
    "IWATER": {
        "PARAMETERS": {"CSNOFG": 1, "RTOPFG": 0, "VRSFG": 0, "VNNFG": 0, "RTLIFG": 1, "LSUR": 200.0, "SLSUR": 0.010,
                       "NSUR": 0.010, "RETSC": 0.01, "PETMAX": 40.0, "PETMIN": 35.0},
        "STATES": {"RETS": 0.01, "SURS": 0.01},
    },
    "SOLIDS": {"PARAMETERS": {"KEIM": 0.08, "JEIM": 1.9, "ACCSDP": 0.01, "REMSDP": 0.5, "SLDS": 0.2}},  # test10.uci:250-260
    "IWTGAS": {"PARAMETERS": {"ELEV": 410.0, "AWTF": 40.0, "BWTF": 0.8, "SDLFAC": 0.0, "SLIFAC": 0.0}},  # :264-268 + readUCI.py:244-250
}

# default SAVE sets, hsp2tools/data/SaveTable.csv
SAVE_DEFAULT = {
    "SNOW": ("COVINX", "DULL", "PACKF", "PACKI", "PACKW", "PACK", "PAKTMP", "RAINF", "RDENPF", "SKYCLR",
             "SNOCOV", "WYIELD", "XLNMLT"),
    "PWATER": ("AGWO", "AGWS", "CEPS", "GWVS", "IFWO", "IFWS", "LZS", "PERO", "SURO", "SURS", "UZS"),
    "IWATER": ("RETS", "SURO", "SURS"),
    "SOLIDS": (), "IWTGAS": (),  # SaveTable.csv rows 256-263: every SOLIDS / IWTGAS series defaults to 0
}


def perlnd():
    return copy.deepcopy(P001)


def implnd():
    return copy.deepcopy(I001)


def iqual_tables():
    """IQUAL of IMPLND 1 (test10.uci:272-290: one constituent, COD, associated with solids and with overland flow) the way
    readUCI lays it out (readUCI.py:690-711 per-constituent tables, :252-258 LAT-FACTOR defaults)."""
    return {
        "PARAMETERS": {"NQUAL": 1, "SDLFAC": 0.0, "SLIFAC": 0.0},
        "IQUAL1_FLAGS": {"QUALID": "COD", "QTYID": "LB", "QSDFG": 1, "VPFWFG": 0, "QSOFG": 1, "VQOFG": 0},
        "IQUAL1_PARAMETERS": {"SQO": 1.20, "POTFW": 0.175, "ACQOP": 0.02, "SQOLIM": 2.0, "WSQOP": 1.7},
    }


def pqual_tables():
    """A three-constituent PQUAL block (the table layout of readUCI.py:690-711; flag combinations after
    tests/GRW_Plaster/HSPF.uci:1385-1560): sediment-associated, overland-flow-associated with subsurface concentrations,
    and QSOFG=2 with atmospheric deposition."""
    return {
        "PARAMETERS": {"NQUAL": 3, "SDLFAC": 0.0, "SLIFAC": 0.3, "ILIFAC": 0.2, "ALIFAC": 0.1},
        "FLAGS": dict({"PQADFG%d" % k: 0 for k in range(1, 21)}, PQADFG5=1, PQADFG6=1),
        "PQUAL1_FLAGS": {"QUALID": "PO4", "QTYID": "LB", "QSDFG": 1, "VPFWFG": 1, "VPFSFG": 0, "QSOFG": 0, "VQOFG": 0,
                         "QIFWFG": 0, "VIQCFG": 0, "QAGWFG": 0, "VAQCFG": 0},
        "PQUAL1_PARAMETERS": {"SQO": 0.0, "POTFW": 1.2, "POTFS": 0.3, "ACQOP": 0.0, "SQOLIM": 0.000001, "WSQOP": 1.64,
                              "IOQC": 0.0, "AOQC": 0.0},
        "PQUAL1_MONTHLY/POTFW": monthly([1.0, 1.0, 1.1, 1.3, 1.5, 1.8, 2.0, 2.0, 1.7, 1.4, 1.2, 1.0]),
        "PQUAL2_FLAGS": {"QUALID": "E.COLI", "QTYID": "CFU", "QSDFG": 0, "VPFWFG": 0, "VPFSFG": 0, "QSOFG": 1, "VQOFG": 1,
                         "QIFWFG": 1, "VIQCFG": 1, "QAGWFG": 1, "VAQCFG": 3},
        "PQUAL2_PARAMETERS": {"SQO": 4.0e8, "POTFW": 0.0, "POTFS": 0.0, "ACQOP": 1.0e6, "SQOLIM": 8.0e8, "WSQOP": 0.7,
                              "IOQC": 2.83e3, "AOQC": 1.42e3},
        "PQUAL2_MONTHLY/ACQOP": monthly([1.0e6, 1.0e6, 1.5e6, 2.0e6, 3.0e6, 4.0e6, 4.0e6, 4.0e6, 3.0e6, 2.0e6, 1.5e6, 1.0e6]),
        "PQUAL2_MONTHLY/SQOLIM": monthly([8.0e8, 8.0e8, 9.0e8, 1.0e9, 1.2e9, 1.4e9, 1.4e9, 1.4e9, 1.2e9, 1.0e9, 9.0e8, 8.0e8]),
        "PQUAL2_MONTHLY/IOQC": monthly([2.0e3, 2.0e3, 2.4e3, 2.8e3, 3.2e3, 3.6e3, 3.6e3, 3.6e3, 3.2e3, 2.8e3, 2.4e3, 2.0e3]),
        "PQUAL2_MONTHLY/AOQC": monthly([10., 10., 11., 12., 13., 14., 14., 14., 13., 12., 11., 10.]),
        "PQUAL3_FLAGS": {"QUALID": "NH3", "QTYID": "LB", "QSDFG": 1, "VPFWFG": 0, "VPFSFG": 0, "QSOFG": 2, "VQOFG": 0,
                         "QIFWFG": 1, "VIQCFG": 0, "QAGWFG": 1, "VAQCFG": 0},
        "PQUAL3_PARAMETERS": {"SQO": 0.6, "POTFW": 0.4, "POTFS": 0.1, "ACQOP": 0.05, "SQOLIM": 1.0, "WSQOP": 1.1,
                              "IOQC": 0.8, "AOQC": 0.3},
        "PQUAL3_MONTHLY/PQADFX": monthly([.001, .001, .002, .002, .003, .003, .003, .003, .002, .002, .001, .001]),
        "PQUAL3_MONTHLY/PQADCN": monthly([.01, .01, .012, .014, .016, .018, .018, .016, .014, .012, .01, .01]),
    }
