"""Host logic of the CTA-pair GEMM launch (csrc/gemm_pair.cu: pick_bn, chaining, build_schedule) through
f2g_gemm_plan -- no device.  The properties the kernel's correctness rests on: every tile runs exactly once,
on every pair all producer tiles come before any consumer tile (a chained launch cannot deadlock while its
CTAs are co-resident), and the consumers' expected counts match how their producers publish."""
import collections

import pytest

from flow2gan_b200 import _lib as L

A0 = 1 << 20          # fake, 16-byte aligned device addresses: the plan never dereferences them


def _mlp(shapes, bn2=None):
    descs, cnt = [], 0x7000000
    for i, (M, C) in enumerate(shapes):
        H = 3 * C
        done = cnt + 4096 * i
        descs.append(L.gemm_desc(A0, A0, A0, M, H, C, C, C, H, bias=A0, slope=A0, act=L.ACT_PRELU, ab_f16=1, c_f16=1,
                                 done_counter=done))
    for i, (M, C) in enumerate(shapes):
        H = 3 * C
        descs.append(L.gemm_desc(A0, A0, A0, M, C, H, H, H, C, bias=A0, res=A0, ld_res=C, res_scale=A0, ab_f16=1,
                                 bn=(bn2 or {}).get(C, 128), wait_counter=cnt + 4096 * i))
    return descs


def _check_cover(plan):
    seen = collections.Counter()
    for lst in plan["lists"]:
        for t in lst:
            seen[t] += 1
    want = 0
    for pi, p in enumerate(plan["problems"]):
        for m in range(p["m_tiles"]):
            for n in range(p["n_tiles"]):
                for k in range(p["split_k"]):
                    assert seen[(pi, m, n, k)] == 1, (pi, m, n, k)
                    want += 1
    assert want == plan["tiles"] == sum(seen.values())


@pytest.mark.parametrize("pairs", [74, 66, 8])
def test_chained_block_launch_plan(pairs):
    shapes = ((1568, 768), (2352, 512), (3136, 384))            # the three branches at the bench shape
    plan = L.gemm_plan(_mlp(shapes, {768: 256, 512: 192, 384: 128}), pairs)
    assert plan["scheduled"] and plan["pairs"] == pairs
    _check_cover(plan)
    P = plan["problems"]
    assert [p["waits"] for p in P] == [0, 0, 0, 1, 1, 1]          # producers first in the tile numbering
    assert [p["K"] for p in P] == [768, 512, 384, 2304, 1536, 1152]   # decreasing K inside a phase
    assert [p["bn"] for p in P[3:]] == [256, 192, 128]           # caller's N tile for the chained consumers
    for p in P[:3]:
        assert p["tma_c"] == 1 and p["bn"] % 64 == 0 and p["publishes"] == 1
    for c in P[3:]:
        prod = next(p for p in P[:3] if p["M"] == c["M"])
        assert c["wait_count"] == 2 * prod["n_tiles"] * 8         # every epilogue warp of both CTAs publishes
    for lst in plan["lists"]:                                     # no consumer tile ahead of a producer tile
        phases = [P[t[0]]["waits"] for t in lst]
        assert phases == sorted(phases), phases
    # longest-processing-time assignment: no pair carries much more than the mean (kernel's own cost model)
    cost = lambda p: -(-p["K"] // 64) * (256 + p["bn"]) + 8 * p["bn"]
    loads = [sum(cost(P[t[0]]) for t in lst) for lst in plan["lists"]]
    assert max(loads) <= 1.25 * sum(loads) / len(loads) + max(cost(p) for p in P)


def test_plain_store_path_and_unchained_counts():
    d = L.gemm_desc(A0, A0, A0, 300, 250, 96, 104, 104, 264, bias=A0, slope=A0, act=L.ACT_PRELU, ab_f16=1, c_f16=1)
    plan = L.gemm_plan([d])
    assert plan["problems"][0]["tma_c"] == 0                      # N % 8 != 0: 16-byte clip granule of a TMA store
    d = L.gemm_desc(A0, A0, A0, 300, 248, 96, 104, 104, 264, bias=A0, slope=A0, act=L.ACT_PRELU, ab_f16=1, c_f16=1)
    plan = L.gemm_plan([d])
    assert plan["problems"][0]["tma_c"] == 1 and plan["problems"][0]["bn"] == 64     # few tiles: narrowed to fill the pairs
    _check_cover(plan)


def test_round_robin_and_oversized_groups():
    d = L.gemm_desc(A0, A0, A0, 256 * 30, 1024, 512, 512, 512, 1024)          # 120 identical tiles
    plan = L.gemm_plan([d], 74)
    _check_cover(plan)
    flat = lambda t: t[1] * plan["problems"][0]["n_tiles"] + t[2]
    for q, lst in enumerate(plan["lists"]):
        assert [flat(t) for t in lst] == list(range(q, 120, 74))
    big = L.gemm_desc(A0, A0, A0, 256 * 300, 1024, 256, 256, 256, 1024)       # 1200 tiles > schedule capacity
    plan = L.gemm_plan([big], 74)
    assert not plan["scheduled"] and plan["tiles"] == 1200


def test_plan_rejects_bad_groups():
    good = L.gemm_desc(A0, A0, A0, 512, 512, 512, 512, 512, 512, ab_f16=1)
    with pytest.raises(RuntimeError, match="wait_counter without a producer"):
        L.gemm_plan([L.gemm_desc(A0, A0, A0, 512, 512, 512, 512, 512, 512, ab_f16=1, wait_counter=0x7000000), good])
    with pytest.raises(RuntimeError, match="16B aligned"):
        L.gemm_plan([L.gemm_desc(A0 + 4, A0, A0, 512, 512, 512, 512, 512, 512)])
