"""The reference's CSV data path (J:372-400; Python/gan.ipynb:104-151): format round trip, batching, one-hot, latent grid."""
import numpy as np
import pytest

from gan_deeplearning4j_b200 import data


def test_csv_round_trip_batches_and_one_hot(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (23, 784)); y = rng.integers(0, 10, 23)
    p = tmp_path / "mnist_train.csv"
    data.write_csv(str(p), x, y)
    first = open(p).readline().strip().split(",")
    assert len(first) == 785 and all(len(t.split(".")[1]) == 2 for t in first)          # "%.2f", label in column 784
    rec = data.read_csv(str(p))
    assert rec.shape == (23, 785)
    it = data.RecordReaderDataSetIterator(rec, 10, 784, 10)
    sizes = []; n = 0
    for f, l in it:
        sizes.append(len(f)); assert f.shape[1] == 784 and l.shape[1] == 10
        np.testing.assert_allclose(f, np.round(x[n:n + len(f)], 2), atol=5e-3)
        assert np.array_equal(l.argmax(1), y[n:n + len(f)]) and np.all(l.sum(1) == 1)
        n += len(f)
    assert sizes == [10, 10, 3]                                                            # ragged tail is kept
    with pytest.raises(ValueError):
        data.RecordReaderDataSetIterator(rec, 10, 785, 10)
    with pytest.raises(ValueError):
        data.RecordReaderDataSetIterator(rec, 10, 784, 5)                                  # labels outside numClasses
    assert data.read_csv(str(p), num_lines_to_skip=3).shape == (20, 785)


def test_latent_grid_matches_J382_389():
    z = data.latent_grid(10)
    assert z.shape == (100, 2) and np.allclose(z[0], [-1, -1]) and np.allclose(z[9], [-1, 1]) and np.allclose(z[10], [-1 + 2 / 9, -1]) and np.allclose(z[-1], [1, 1])
