"""MicroBatcher: concurrent single-sample calls are served by one batched call,
every caller gets its own result, a lone sequential caller is not delayed, errors reach every caller."""
import threading
import time

import pytest

from multimodal_detection_consistency_b200.batching import MicroBatcher


def _run_threads(n, fn):
    out, errs = [None] * n, [None] * n
    start = threading.Barrier(n)

    def work(i):
        start.wait()
        try:
            out[i] = fn(i)
        except Exception as e:  # noqa: BLE001
            errs[i] = e
    ts = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out, errs


def test_concurrent_calls_share_rounds_and_get_their_own_results():
    calls = []

    def batch_fn(key, items):
        calls.append((key, list(items)))
        time.sleep(0.005)                       # a "launch"
        return [(key, x * x) for x in items]
    mb = MicroBatcher(batch_fn, max_batch=64, max_delay_s=0.05)
    out, errs = _run_threads(16, lambda i: mb.submit(i, key=i % 2))
    assert errs == [None] * 16
    assert out == [(i % 2, i * i) for i in range(16)]
    assert all(len({k}) == 1 for k, _ in calls)
    assert len(calls) < 16 and sum(len(it) for _, it in calls) == 16
    assert mb.stats()["items"] == 16 and mb.stats()["mean_batch"] > 1.0


def test_max_batch_bounds_a_round():
    sizes = []

    def batch_fn(key, items):
        sizes.append(len(items))
        return list(items)
    mb = MicroBatcher(batch_fn, max_batch=4, max_delay_s=0.2)
    out, errs = _run_threads(16, lambda i: mb.submit(i))
    assert errs == [None] * 16 and sorted(out) == list(range(16))
    assert max(sizes) <= 4 and sum(sizes) == 16 and 4 in sizes     # full rounds close without waiting


def test_sequential_caller_is_not_delayed():
    mb = MicroBatcher(lambda k, items: [x + 1 for x in items], max_delay_s=0.5, idle_s=0.01)
    assert mb.submit(1) == 2
    time.sleep(0.05)
    t0 = time.monotonic()
    for i in range(20):
        assert mb.submit(i) == i + 1
    assert time.monotonic() - t0 < 0.4          # 20 calls, none waited for followers
    assert mb.stats()["rounds"] == 21


def test_errors_reach_every_caller_of_the_round():
    def boom(key, items):
        raise RuntimeError("encoder down")
    mb = MicroBatcher(boom, max_delay_s=0.05)
    out, errs = _run_threads(4, lambda i: mb.submit(i))
    assert out == [None] * 4 and all(isinstance(e, RuntimeError) for e in errs)
    with pytest.raises(RuntimeError):
        MicroBatcher(lambda k, items: [], max_delay_s=0.0).submit(1)     # wrong result count
