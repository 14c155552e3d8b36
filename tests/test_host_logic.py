"""Host-side helpers of the end-to-end path that need no GPU (phasegen/hostmem.py)."""


def test_gpu_locality_parsing_and_best_effort_binding(monkeypatch):
    """phasegen.hostmem: cpulist parsing and the best-effort contract (no topology -> nothing changes, no exception)."""
    import os
    from phasegen import hostmem
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    monkeypatch.setattr(hostmem, "gpu_locality", lambda i: (None, None))
    assert hostmem.bind_to_gpu_node(0) == {"numa_node": None}
    assert os.sched_getaffinity(0) == before
    # a node whose CPUs this process may not use: reported, not applied
    monkeypatch.setattr(hostmem, "gpu_locality", lambda i: (1, {10 ** 6}))
    r = hostmem.bind_to_gpu_node(0)
    assert r["numa_node"] == 1 and r["bound"] is False and os.sched_getaffinity(0) == before
    # a node that covers the current mask: bound to the intersection
    monkeypatch.setattr(hostmem, "gpu_locality", lambda i: (0, set(before) | {10 ** 6}))
    r = hostmem.bind_to_gpu_node(0)
    assert r == {"numa_node": 0, "cpus": len(before), "bound": True} and os.sched_getaffinity(0) == before


def test_ramped_sub_batch_sizes():
    """pipeline.ramp_sizes: wave-aligned sub-batches that ramp 1-2-4 waves up, `max_waves` in the middle, 4-2-1 down, always adding
    up to the batch; short batches drop the outer steps instead of producing slivers."""
    from phasegen.pipeline import ramp_sizes
    cap = lambda w: max(2, 2 * ((w * 74) // 8))                  # the BASELINE shape: 18 clips per wave of the heaviest convolution
    assert ramp_sizes(256, cap, 2) == [18, 36, 74, 74, 36, 18]
    for B in (7, 20, 36, 40, 64, 128, 256, 300, 512, 1000):
        sizes = ramp_sizes(B, cap, 2)
        assert sum(sizes) == B and min(sizes) >= 1
        if len(sizes) >= 3:
            assert sizes[0] == sizes[-1] == 18                     # one-wave head and tail: the only exposed copies
            assert all(b <= 2 * a + 2 or b <= 110 for a, b in zip(sizes, sizes[1:3]))   # each upload fits under its predecessor's compute
    assert ramp_sizes(20, cap, 2) == [20] and ramp_sizes(36, cap, 2) == [36]
