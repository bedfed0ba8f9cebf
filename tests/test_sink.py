"""The product results store (hspsquared_b200/sink.py NpyResults: SupportsWriteTS / SupportsReadTS, hsp2io/protocols.py:26-43)
and the streaming writer: a batch streamed into it chunk by chunk must read back, per (segment, activity), as exactly the
float32 frame the reference's save_timeseries hands IOManager.write_ts (utilities.py:292-333), here built by results.save_batch."""
import numpy as np
import pandas as pd
import pytest

from hspsquared_b200 import synth, test10
from hspsquared_b200.calendar import Calendar
from hspsquared_b200.engine import LandBatch
from hspsquared_b200.results import save_batch
from hspsquared_b200.sink import NpyResults, stream_batch

from . import parity


@pytest.fixture(autouse=True)
def _use_emu(emu_library):
    yield


def test_stream_batch_equals_save_timeseries_frames(tmp_path):
    parity.check_streaming_sink(tmp_path, n=37, days=9, chunk=50)


def test_single_frames_round_trip_and_missing_tables_read_empty(tmp_path):
    sink = NpyResults(tmp_path / "r")
    t = pd.date_range("1976-01-01", periods=30, freq="60min")[1:]
    df = pd.DataFrame(np.random.default_rng(1).random((29, 3)).astype(np.float32), index=t, columns=["AGWO", "IFWO", "SURO"])
    sink.write_ts(df, "RESULT", "PERLND", "P001", "PWATER", compress=True)
    sink.close()
    again = NpyResults(tmp_path / "r")
    g = again.read_ts("RESULT", "PERLND", "P001", "PWATER")
    assert g.equals(df) and g.to_numpy().dtype == np.float32
    assert again.read_ts("RESULT", "PERLND", "P002", "PWATER").empty  # hdf.py:91-92
