"""Shared helpers of the parity tests."""
import json
import os

import numpy as np

from oracle import pyoracle as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_search.json")

# north_star tolerance: ids/order exact except among results whose distances are within
# 1e-5 relative of each other (vs the reference's float64 distances)
TIE_RTOL = 1e-5
# fp64-verified distances: every operation is IEEE-exact and identically ordered on both sides, including Go's
# math.Acos, which oracle and device both restate (Cephes algorithm) instead of calling their libm: the returned
# distances are BIT-IDENTICAL to the oracle's, for every quantization and both metrics
DIST_RTOL = 0.0


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def golden_case_inputs(case):
    n, dims, bits, seed = case["n"], case["dims"], case["bits"], case["seed"]
    codes = o.synth_rows(seed, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64) * 3 + 1
    queries = o.synth_queries(seed + 1, 0, 3, dims)
    passmask = None if not case["filter_mod"] else (ids % case["filter_mod"] == 0).astype(np.uint8)
    return codes, ids, queries, passmask


def assert_results_match(got_ids, got_dist, ref_ids, ref_dist, true_dist_of=None, what=""):
    """Ordered result lists agree: same length, distances equal to DIST_RTOL, ids equal except inside
    groups of reference distances closer than TIE_RTOL (where any order / boundary member is allowed,
    provided the returned id really has that distance)."""
    got_ids, ref_ids = np.asarray(got_ids), np.asarray(ref_ids)
    got_dist, ref_dist = np.asarray(got_dist, dtype=np.float64), np.asarray(ref_dist, dtype=np.float64)
    assert got_ids.shape == ref_ids.shape, f"{what}: {got_ids.size} results, reference has {ref_ids.size}"
    if ref_ids.size == 0:
        return
    scale = np.maximum(np.abs(ref_dist), 1e-300)
    assert np.all(np.abs(got_dist - ref_dist) <= TIE_RTOL * scale + 1e-300), \
        f"{what}: distances differ beyond tolerance\n got {got_dist}\n ref {ref_dist}"
    assert np.all(np.diff(got_dist) >= 0), f"{what}: output not ascending: {got_dist}"
    for i in range(ref_ids.size):
        if got_ids[i] == ref_ids[i]:
            assert abs(got_dist[i] - ref_dist[i]) <= DIST_RTOL * scale[i] + 1e-300, \
                f"{what}: id {got_ids[i]} distance {got_dist[i]!r} vs reference {ref_dist[i]!r}"
            continue
        # different id at this rank: only legal inside a tie group
        near = np.abs(ref_dist - ref_dist[i]) <= TIE_RTOL * scale[i]
        in_group = got_ids[i] in set(ref_ids[near].tolist())
        boundary = near[-1]  # tie group reaches the k-th place: an outside member may replace one inside
        assert in_group or boundary, \
            f"{what}: rank {i}: got id {got_ids[i]} (d={got_dist[i]!r}), reference id {ref_ids[i]} (d={ref_dist[i]!r})"
        if true_dist_of is not None:
            td = true_dist_of(int(got_ids[i]))
            assert abs(td - got_dist[i]) <= DIST_RTOL * max(abs(td), 1e-300) + 1e-300, \
                f"{what}: id {got_ids[i]} reported d={got_dist[i]!r} but its distance is {td!r}"


def assert_radius_match(got_ids, got_dist, ref_ids, ref_dist, radius, what=""):
    """Radius results: same membership except records whose distance is within TIE_RTOL of the radius."""
    got = dict(zip(np.asarray(got_ids).tolist(), np.asarray(got_dist).tolist()))
    ref = dict(zip(np.asarray(ref_ids).tolist(), np.asarray(ref_dist).tolist()))
    for i, d in ref.items():
        if i not in got:
            assert abs(d - radius) <= TIE_RTOL * radius, f"{what}: id {i} (d={d!r}) missing from radius result"
        else:
            assert abs(got[i] - d) <= DIST_RTOL * max(abs(d), 1e-300) + 1e-300, f"{what}: id {i}: {got[i]!r} vs {d!r}"
    for i, d in got.items():
        assert d <= radius, f"{what}: id {i} has distance {d!r} > radius {radius!r}"
        if i not in ref:
            assert abs(d - radius) <= TIE_RTOL * radius, f"{what}: unexpected id {i} (d={d!r})"
    gd = np.asarray(got_dist)
    assert np.all(np.diff(gd) >= 0), f"{what}: radius output not ascending"
