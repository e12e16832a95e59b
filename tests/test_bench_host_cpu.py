"""CPU: host-side helpers of bench.py that run before any GPU work."""
import os

import bench


def test_numa_binding_degrades_without_nvidia_smi():
    """No nvidia-smi / sysfs entry (this container): the rank stays unbound, says why, and its CPU affinity is untouched."""
    before = os.sched_getaffinity(0)
    info = bench.bind_to_gpu_numa_node(0)
    assert set(info) >= {"node", "cpus"}
    if info["node"] is None:
        assert os.sched_getaffinity(0) == before
    else:  # on a GPU box: bound to a non-empty subset
        assert 0 < info["cpus"] <= len(before) and os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)


def test_kernel_sources_hash_matches_the_committed_capture():
    """profiles/traffic.json belongs to the render kernels' sources as committed: bench.py reports `roofline.traffic` only then."""
    import json

    import pytest

    tj = json.load(open(os.path.join(bench.ROOT, "profiles", "traffic.json")))
    if tj["kernel_sources_sha"] != bench.kernel_sources_sha():
        pytest.skip("the render kernels' sources changed since the last ncu capture: bench.py will report roofline.traffic = null "
                    "until profiles/traffic.json is regenerated (tools/make_traffic_json.py)")
    assert tj["dram_bytes_per_launch"] > 1e9 and set(tj["limiter"]) >= {"sample_issue_active_pct", "shade_issue_active_pct"}
