"""Host logic of the GPU-resident replay pool, exercised on CPU tensors:
the class is plain tensor indexing, so the same code path runs on ``cuda`` devices."""
import random

import torch

from adaptiveisp_b200.config import make_cfg
from adaptiveisp_b200.replay_pool import DeviceReplayPool, STATE_STEP_DIM, STATE_STOPPED_DIM


def make_pool(capacity=12, seed=0):
    cfg = make_cfg()
    cfg.replay_memory_size = capacity
    cfg.maximum_trajectory_length = 7
    cfg.over_length_keep_prob = 0.5
    counter = {"n": 0}

    def fetch(n):
        ids = list(range(counter["n"], counter["n"] + n))
        counter["n"] += n
        imgs = torch.stack([torch.full((3, 4, 4), float(i)) for i in ids])
        return imgs, [{"id": i} for i in ids]

    return DeviceReplayPool(cfg, (3, 4, 4), "cpu", fetch, fetch_batch=5, rng=random.Random(seed)), counter, cfg


def test_fill_and_draw_without_replacement():
    pool, counter, cfg = make_pool()
    assert len(pool) == 12 and counter["n"] == 12
    b = pool.get_batch(5)
    assert b.images.shape == (5, 3, 4, 4) and b.states.shape == (5, cfg.num_state_dim)
    assert len(set(b.slots)) == 5 and len(pool) == 7
    for img, meta in zip(b.images, b.meta):            # pixel data and host records stay paired
        assert float(img[0, 0, 0]) == float(meta["id"])
    assert float(b.states.abs().sum()) == 0.0          # fresh records start from the zero state


def test_put_back_keeps_pixels_on_device_and_refills():
    pool, counter, cfg = make_pool()
    b = pool.get_batch(4)
    new_states = b.states.clone()
    new_states[:, STATE_STEP_DIM] = 1.0
    kept = pool.put_back(b.slots, b.images + 0.5, new_states)
    assert kept == 4 and len(pool) == 12 and counter["n"] == 12     # nothing new had to be fetched
    for s, m in zip(b.slots, b.meta):
        assert float(pool.images[s, 0, 0, 0]) == m["id"] + 0.5
        assert float(pool.states[s, STATE_STEP_DIM]) == 1.0 and pool.meta[s] is m


def test_finished_records_are_dropped_when_met_and_overlong_ones_by_coin_flip():
    pool, counter, cfg = make_pool(seed=3)
    b = pool.get_batch(6)
    st = b.states.clone()
    st[:3, STATE_STOPPED_DIM] = 1.0                      # three finished trajectories
    st[:, STATE_STEP_DIM] = 2.0
    pool.put_back(b.slots, b.images, st)
    finished = set(b.slots[:3])
    seen = set()
    for _ in range(6):
        nb = pool.get_batch(4)
        assert not (set(nb.slots) & finished) or all(pool._stopped[s] != 1 for s in nb.slots)
        seen.update(m["id"] for m in nb.meta)
        pool.put_back(nb.slots, nb.images, nb.states)
    assert counter["n"] > 12                             # the dropped records were replaced by fresh ones
    # over-long trajectories: kept with probability over_length_keep_prob
    pool, counter, cfg = make_pool(capacity=40, seed=5)
    b = pool.get_batch(40)
    st = b.states.clone()
    st[:, STATE_STEP_DIM] = 9.0                          # > maximum_trajectory_length
    kept = pool.put_back(b.slots, b.images, st)
    assert 8 <= kept <= 32 and len(pool) == 40 and counter["n"] == 40 + (40 - kept)


def test_discard_bad_batch_refills():
    pool, counter, cfg = make_pool()
    b = pool.get_batch(3)
    pool.discard(b.slots)
    assert len(pool) == 12 and counter["n"] == 15
    assert pool.average_trajectory() == 0.0
