"""CPU: trajectory part files — ``PartCacheWriter`` writes the reference's on-disk format
(quantrl/io.py:56-80) on a background thread and ``load_parts`` stitches a batch back into the reference env's ``data``
cache (fixture G3)."""
import os

import numpy as np
import pytest

import em_oracle as eo


@pytest.fixture(scope="module")
def g_em(golden_dir):
    return np.load(os.path.join(golden_dir, "em_linear_env.npz"))


def test_part_files_round_trip_to_the_reference_data_cache(g_em, tmp_path):
    """PartCacheWriter (host side of §8 a10 / f1): the packed blocks of an episode written as the reference's part files
    (io.py:56-80 naming, key ``arr_0``) and stitched by ``load_parts`` reproduce the reference env's ``data`` cache,
    including the row shared by adjacent blocks (base.py:1370-1372)."""
    from quantrl_b200.envs.base import PartCacheWriter
    g, tag = g_em, "n4"
    K, S, ssz = int(g[f"{tag}_K"]), int(g[f"{tag}_S"]), float(g[f"{tag}_ssz"])
    data = g[f"{tag}_data"]
    E = data.shape[0]
    T = np.arange(S + 1) * ssz
    Obs = np.transpose(g[f"{tag}_Observations"], (1, 0, 2))
    io = PartCacheWriter(str(tmp_path / "cache"), cache_dump_interval=100, compressed=True)
    for a in range(S // K):
        t0 = a * K
        io.dump_part_async(eo.pack_rows(T[t0:t0 + K + 1], g[f"{tag}_actions"][:, a], Obs[t0:t0 + K + 1], None,
                                        data[:, t0 + K, -1]), batch_idx=3, part_idx=a)
    io.flush()
    assert sorted(os.listdir(tmp_path / "cache")) == sorted(f"300_399_{a}.npz" for a in range(S // K))
    got = io.load_parts(3)
    assert got.shape == data.shape
    np.testing.assert_allclose(got[:, :, :6], data[:, :, :6], rtol=0, atol=1e-15)
    # cumulative reward column: every row of block a carries the total after block a (base.py:1342-1349); the shared
    # row keeps the later block's value
    for a in range(S // K):
        np.testing.assert_array_equal(got[:, a * K:(a + 1) * K, -1], np.repeat(data[:, (a + 1) * K, -1:], K, 1))
    np.testing.assert_array_equal(io.load_parts(3, idxs=[-1]), got[:, :, -1:])
    with pytest.raises(FileNotFoundError):
        io.load_parts(4)
    io.close()


def test_part_file_write_errors_surface_in_the_callers_thread(tmp_path, monkeypatch):
    """A failing background write is not lost and does not dead-lock the producer: it is raised by the next
    ``dump_part_async`` / ``flush`` / ``close``, and the writer keeps working afterwards."""
    from quantrl_b200.envs.base import PartCacheWriter
    real = np.savez_compressed

    def boom(path, data):
        raise OSError("disk full")

    io = PartCacheWriter(str(tmp_path / "cache"), cache_dump_interval=4)
    monkeypatch.setattr(np, "savez_compressed", boom)
    raised = 0
    for part in range(20):                           # more items than the queue holds: the producer must never block
        try:
            io.dump_part_async(np.zeros((4, 3, 2)), batch_idx=0, part_idx=part)
        except RuntimeError as e:
            raised += 1
            assert "part file" in str(e) and isinstance(e.__cause__, OSError)
    try:
        io.flush()
    except RuntimeError:
        raised += 1
    assert raised >= 1
    monkeypatch.setattr(np, "savez_compressed", real)
    io.dump_part_async(np.ones((4, 3, 2)), batch_idx=1, part_idx=0)
    io.close()
    assert np.array_equal(io.load_parts(1), np.ones((4, 3, 2)))


def test_part_files_of_two_shards_sharing_a_directory_do_not_collide(tmp_path):
    """Sharded batch (``make_sharded_env``: every rank shares ``dir_prefix`` and ``batch_idx``): the part names carry
    GLOBAL env indices, so rank 1 never overwrites rank 0's parts and each rank reads back its own shard."""
    from quantrl_b200.envs.base import PartCacheWriter
    E_total, bounds = 10, [(0, 6), (6, 4)]
    writers = [PartCacheWriter(str(tmp_path / "cache"), cache_dump_interval=cnt, env_offset=off, n_envs_total=E_total)
               for off, cnt in bounds]
    blocks = [np.full((cnt, 3, 2), float(r)) + np.arange(3)[None, :, None] for r, (_, cnt) in enumerate(bounds)]
    for part in range(2):
        for w, b in zip(writers, blocks):
            w.dump_part_async(b + 10 * part, batch_idx=2, part_idx=part)
    for w in writers:
        w.flush()
    assert sorted(os.listdir(tmp_path / "cache")) == ["20_25_0.npz", "20_25_1.npz", "26_29_0.npz", "26_29_1.npz"]
    for r, w in enumerate(writers):
        got = w.load_parts(2)
        assert got.shape == (bounds[r][1], 5, 2) and got[0, 0, 0] == float(r) and got[0, -1, 0] == float(r) + 12
        w.close()
