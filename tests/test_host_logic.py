"""Host-side logic that needs no GPU: the SerialIterator stand-in's epoch wrap-around."""
import numpy as np


def test_array_iterator_completes_last_batch_from_next_epoch():
    """chainer.iterators.SerialIterator(dataset, batch_size, repeat=True, shuffle=True) (srgan_train.py:159-165):
    every batch has batch_size items; a batch reaching the end of the data is completed from the head of the next
    epoch's freshly drawn order, and ``epoch`` advances with that batch."""
    from deepbedmap_b200.train import ArrayIterator
    n, bs = 10, 4
    arrays = {"X": np.arange(n, dtype=np.float32).reshape(n, 1), "Y": np.arange(n, dtype=np.float32)[:, None] * 2}
    it = ArrayIterator(arrays, bs, shuffle=True, seed=3)
    rng = np.random.RandomState(3)
    order0, order1 = rng.permutation(n), rng.permutation(n)
    seen = []
    epochs = []
    for _ in range(5):
        b = it.next()
        assert b["X"].shape == (bs, 1) and np.array_equal(b["Y"], 2 * b["X"])
        seen += [int(v) for v in b["X"][:, 0]]
        epochs.append(it.epoch)
    assert epochs == [0, 0, 1, 1, 2]
    assert seen == list(order0) + list(order1)
    # unshuffled, batch size dividing the data: the epoch flips exactly at the end, nothing is carried over
    it = ArrayIterator(arrays, 5, shuffle=False)
    a, b, c = it.next(), it.next(), it.next()
    assert list(a["X"][:, 0]) == [0, 1, 2, 3, 4] and list(b["X"][:, 0]) == [5, 6, 7, 8, 9] and it.epoch == 1
    assert list(c["X"][:, 0]) == [0, 1, 2, 3, 4]


def test_s2d_tap_mask_matches_the_embedding():
    """flat.s2d_tap_mask: the taps of the 3x3-over-phases embedding of a 4x4 stride-2 filter that touch a real filter
    element (csrc/umma_conv3x3.cu pack mode 2: filter row 2 (ky - 1) + py + 1 in 0..3) -- 4 per phase, their union for
    channel chunks that span phases."""
    from deepbedmap_b200.flat import s2d_tap_mask
    C = 64
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        want = 0
        for tap in range(9):
            ky, kx = 2 * (tap // 3 - 1) + py + 1, 2 * (tap % 3 - 1) + px + 1
            if 0 <= ky <= 3 and 0 <= kx <= 3:
                want |= 1 << tap
        assert bin(want).count("1") == 4
        assert s2d_tap_mask(ph * C, C, C) == want
        assert s2d_tap_mask(ph * C + 16, 32, C) == want
    assert s2d_tap_mask(0, 128, C) == s2d_tap_mask(0, 64, C) | s2d_tap_mask(64, 64, C)
    assert s2d_tap_mask(0, 4 * C, C) == 0x1FF
