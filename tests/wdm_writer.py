"""Minimal WDM writer for tests: one label record per dataset followed by its chain of data records.  Enough of the layout
for hspsquared_b200.wdm.WdmFile to be exercised without the reference's files (which do not travel with the repo)."""
import numpy as np

RECORD = 512


def _date_word(t):
    y, m, d = [int(x) for x in str(t.astype("datetime64[D]")).split("-")]
    h = int((t - t.astype("datetime64[D]")) / np.timedelta64(1, "h"))
    return (y << 14) | (m << 10) | (d << 5) | h


def write_wdm(path, datasets):
    """datasets: list of dicts {dsn, tcode, tstep, tgroup, groups: [(start datetime64[s], [(values, compressed), ...])]}
    -- every block of a group uses the dataset's tcode / tstep."""
    recs = [np.zeros(RECORD, dtype=np.int32)]  # file definition record
    for ds in datasets:
        label = np.zeros(RECORD, dtype=np.int32)
        recs.append(label)
        label[0] = 1  # non-zero link words mark the record as in use
        label[4], label[5] = ds["dsn"], 1
        psa, pdat = 20, 60
        attrs = [(17, ds["tcode"]), (33, ds["tstep"]), (34, ds["tgroup"])]
        label[9], label[10], label[11] = psa, pdat, pdat + 2 + len(ds["groups"]) + 1
        label[psa - 1] = len(attrs)
        for k, (aid, val) in enumerate(attrs):
            label[psa + 1 + 2 * k], label[psa + 2 + 2 * k] = aid, 40 + k + 1
            label[40 + k] = val
        # data records: four link words, then group after group.  A block never spans records: it is split so that it ends
        # at the end of the record, and the chain goes on after the link words of the next one
        state = {"cur": np.zeros(RECORD, dtype=np.int32), "no": len(recs), "off": 4}
        recs.append(state["cur"])

        def hop():
            nxt = np.zeros(RECORD, dtype=np.int32)
            state["cur"][3] = len(recs) + 1
            recs.append(nxt)
            state.update(cur=nxt, no=len(recs) - 1, off=4)

        for g, (start, blocks) in enumerate(ds["groups"]):
            if RECORD - state["off"] < 3:
                hop()
            label[pdat + 1 + g] = ((state["no"] + 1) << 9) | (state["off"] + 1)
            state["cur"][state["off"]] = _date_word(start)
            state["off"] += 1
            for values, compressed in blocks:
                values = np.asarray(values, dtype=np.float32)
                while len(values):
                    if RECORD - state["off"] < 2:
                        hop()
                    cur, off = state["cur"], state["off"]
                    take = len(values) if compressed else min(len(values), RECORD - off - 1)
                    cur[off] = (take << 16) | (ds["tstep"] << 10) | (ds["tcode"] << 7) | ((1 if compressed else 0) << 5)
                    if compressed:
                        cur[off + 1:off + 2] = values[:1].view(np.int32)
                        state["off"] = off + 2
                    else:
                        cur[off + 1:off + 1 + take] = values[:take].view(np.int32)
                        state["off"] = off + 1 + take
                    values = values[take:]
                    if state["off"] >= RECORD - 1:
                        hop()
    head = recs[0]
    head[0], head[28], head[31] = -998, len(recs), len(datasets)
    np.concatenate(recs).astype("<i4").tofile(path)
