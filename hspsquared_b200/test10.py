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
