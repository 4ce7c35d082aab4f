"""Property-based tests (hypothesis) of the pure-logic components against torch's implementations as oracles:
bucket planner, DistributedSampler, key-value store semantics."""
import torch
import torch.distributed as tdist
from hypothesis import given, settings
from hypothesis import strategies as st

import pytorch_distributed_train_b200 as pdt
from pytorch_distributed_train_b200 import _C

_DTYPES = [torch.float32, torch.float16, torch.float64]


@settings(max_examples=60, deadline=None)
@given(sizes=st.lists(st.integers(1, 5000), min_size=1, max_size=40),
       dtype_ix=st.lists(st.integers(0, 2), min_size=40, max_size=40),
       first=st.integers(256, 20000), cap=st.integers(1000, 80000), reverse=st.booleans())
def test_bucket_planner_equals_torch_and_partitions(sizes, dtype_ix, first, cap, reverse):
    params = [torch.empty(n, dtype=_DTYPES[dtype_ix[i]]) for i, n in enumerate(sizes)]
    nbytes = [p.numel() * p.element_size() for p in params]
    keys = [hash(str(p.dtype)) & 0xFFFF for p in params]
    order = list(reversed(range(len(params)))) if reverse else []
    ours, _ = _C.plan_buckets(nbytes, keys, [first, cap], order)
    if reverse:
        ref, _ = tdist._compute_bucket_assignment_by_size([params[i] for i in order], [first, cap], [False] * len(params), order)
    else:
        ref, _ = tdist._compute_bucket_assignment_by_size(params, [first, cap])
    if reverse and len({p.dtype for p in params}) > 1:
        # with an explicit ready order torch does not sort, and emits the per-dtype leftover buckets in the iteration
        # order of a std::unordered_map — only the bucket *contents* are a contract there (rank 0's layout is
        # broadcast to everybody anyway, C6)
        assert sorted(ours) == sorted(ref)
    else:
        assert ours == ref
    flat = [i for b in ours for i in b]
    assert sorted(flat) == list(range(len(params)))                       # a partition
    for b in ours:
        assert len({params[i].dtype for i in b}) == 1                     # never mixes dtypes


@settings(max_examples=80, deadline=None)
@given(n=st.integers(1, 300), replicas=st.integers(1, 9), shuffle=st.booleans(), drop_last=st.booleans(),
       seed=st.integers(0, 2 ** 31 - 1), epoch=st.integers(0, 50))
def test_distributed_sampler_equals_torch(n, replicas, shuffle, drop_last, seed, epoch):
    from torch.utils.data.distributed import DistributedSampler as TorchSampler

    data = list(range(n))
    seen = []
    for rank in range(replicas):
        a = pdt.DistributedSampler(data, num_replicas=replicas, rank=rank, shuffle=shuffle, seed=seed, drop_last=drop_last)
        b = TorchSampler(data, num_replicas=replicas, rank=rank, shuffle=shuffle, seed=seed, drop_last=drop_last)
        a.set_epoch(epoch)
        b.set_epoch(epoch)
        ia = list(iter(a))
        assert ia == list(iter(b)) and len(a) == len(b) == len(ia)
        seen += ia
    if not drop_last:
        assert set(seen) == set(data)                                      # every sample is visited (some twice: wrap padding)


_KEY = st.sampled_from(["a", "b", "c/d", ""])
_OPS = st.lists(st.one_of(
    st.tuples(st.just("set"), _KEY, st.binary(max_size=12)),
    st.tuples(st.just("add"), _KEY, st.integers(-5, 5)),
    st.tuples(st.just("append"), _KEY, st.binary(max_size=6)),
    st.tuples(st.just("delete"), _KEY, st.none()),
    st.tuples(st.just("cas"), _KEY, st.tuples(st.binary(max_size=3), st.binary(max_size=3))),
), max_size=40)


@settings(max_examples=60, deadline=None)
@given(ops=_OPS)
def test_hash_store_behaves_like_a_dict_model(ops):
    """set / add / append / delete_key / compare_set against a plain-dict model (HashStore shares KVState with the
    TCP store's server side)."""
    s = _C.HashStore()
    model = {}
    for op, k, v in ops:
        if op == "set":
            s.set(k, v)
            model[k] = bytes(v)
        elif op == "add":
            cur = model.get(k)
            if cur is not None:
                try:
                    int(cur.decode())
                except (ValueError, UnicodeDecodeError):
                    continue  # add on a non-numeric value is an error in both worlds; not the property under test
            got = s.add(k, v)
            new = (int(cur.decode()) if cur else 0) + v
            model[k] = str(new).encode()
            assert got == new
        elif op == "append":
            s.append(k, v)
            model[k] = model.get(k, b"") + bytes(v)
        elif op == "delete":
            assert s.delete_key(k) == (k in model)
            model.pop(k, None)
        elif op == "cas":
            expected, desired = v
            got = bytes(s.compare_set(k, expected, desired))
            if k not in model:
                if expected == b"":
                    model[k] = bytes(desired)
                    assert got == bytes(desired)
                else:
                    assert got == bytes(expected)       # torch semantics: missing key + non-empty expectation → echo `expected`
            elif model[k] == bytes(expected):
                model[k] = bytes(desired)
                assert got == bytes(desired)
            else:
                assert got == model[k]
    assert s.num_keys() == len(model)
    for k, v in model.items():
        assert s.check([k]) and bytes(s.get(k)) == v
