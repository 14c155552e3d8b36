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
