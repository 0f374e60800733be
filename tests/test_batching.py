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


def test_a_failing_item_does_not_fail_its_round():
    """one bad sample inside a coalesced round must only fail its own caller (the reference isolates
    failures per sample); the leader re-runs the round item by item."""
    import threading
    from multimodal_detection_consistency_b200.batching import MicroBatcher
    calls = []

    def fn(key, items):
        calls.append(list(items))
        if any(x == "bad" for x in items):
            raise ValueError("corrupt sample")
        return [x * 2 for x in items]

    mb = MicroBatcher(fn, max_batch=4, max_delay_s=0.5, idle_s=10.0)
    items = [1, "bad", 3, 4]
    out, errs = {}, {}
    gate = threading.Barrier(4)

    def work(x):
        gate.wait()
        try:
            out[x] = mb.submit(x)
        except ValueError as e:
            errs[x] = e

    th = [threading.Thread(target=work, args=(x,)) for x in items]
    [t.start() for t in th]
    [t.join() for t in th]
    assert out == {1: 2, 3: 6, 4: 8}
    assert list(errs) == ["bad"]
    assert mb.isolated <= 1 and max(len(c) for c in calls) >= 1
    # a lone failing call still raises its own error
    try:
        mb.submit("bad")
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
