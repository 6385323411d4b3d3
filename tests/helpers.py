"""Shared helpers for the parity tests."""
import numpy as np

STATE_INT_KEYS = ["states", "ids", "hits", "ages", "last_frame", "active", "row_assign"]
STATE_FLOAT_KEYS = ["poses", "vel", "scores", "predicted", "centers"]


def compare_state(got: dict, ref: dict, D: int, T: int, where="") -> list:
    """Bit-exact comparison of a tracker state dump; returns a list of mismatch strings.
    Entries the reference leaves undefined are masked: per-slot values of slots that were
    never used keep their initial zeros on both sides, so no masking is needed there;
    col_assign is only defined for [0, D) and cost for the flat region [0, T*D)."""
    bad = []
    for k in STATE_INT_KEYS:
        if not np.array_equal(got[k], ref[k]):
            idx = np.nonzero(got[k] != ref[k])[0][:5]
            bad.append(f"{where} {k} differs at {idx}: got {got[k][idx]} ref {ref[k][idx]}")
    for k in STATE_FLOAT_KEYS:
        if got[k].tobytes() != ref[k].tobytes():
            d = np.nonzero(got[k].ravel().view(np.uint32) != ref[k].ravel().view(np.uint32))[0][:5]
            bad.append(f"{where} {k} differs at flat {d}: got {got[k].ravel()[d]} ref {ref[k].ravel()[d]}")
    if not np.array_equal(got["col_assign"][:D], ref["col_assign"][:D]):
        bad.append(f"{where} col_assign differs: got {got['col_assign'][:D]} ref {ref['col_assign'][:D]}")
    if got["cost"][: T * D].tobytes() != ref["cost"][: T * D].tobytes():
        d = np.nonzero(got["cost"][: T * D].view(np.uint32) != ref["cost"][: T * D].view(np.uint32))[0][:5]
        bad.append(f"{where} cost differs at flat {d}: got {got['cost'][d]} ref {ref['cost'][d]}")
    if not np.array_equal(got["scalars"], ref["scalars"]):
        bad.append(f"{where} scalars differ: got {got['scalars']} ref {ref['scalars']}")
    return bad


def valid_records(out, counts) -> bytes:
    """The TrackOutput records a handle reports ([B, Dm] buffer + counts): only the first counts[b] of every stream are defined
    (pipelined handles keep one record buffer per ring slot, so what lies behind them differs from a serial handle's)."""
    return b"".join(out[b, : counts[b]].tobytes() for b in range(len(counts)))
