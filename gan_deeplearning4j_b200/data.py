"""The reference's data path: CSV files written by the notebook (Python/gan.ipynb:104-151 -- 784 pixel columns
formatted "%.2f" plus the class label in column 784, comma separated, no header) read the way the driver does:
CSVRecordReader(numLinesToSkip, ",") + RecordReaderDataSetIterator(reader, batchSize, labelIndex=784, numClasses=10) (J:372-377,395-400),
and the 10x10 latent grid of J:382-389.  Host-side only; nothing here touches the GPU.
"""
from __future__ import annotations

from typing import Iterator, Tuple

import numpy as np


def read_csv(path: str, num_lines_to_skip: int = 0, delimiter: str = ",") -> np.ndarray:
    """CSVRecordReader(numLinesToSkip, delimiter): every remaining line is one record of doubles."""
    return np.loadtxt(path, delimiter=delimiter, skiprows=num_lines_to_skip, dtype=np.float32, ndmin=2)


class RecordReaderDataSetIterator:
    """RecordReaderDataSetIterator(reader, batchSize, labelIndex, numClasses): features = every column except `labelIndex`,
    labels = one-hot of that column; the last minibatch may be smaller (DL4J does not drop it)."""

    def __init__(self, records: np.ndarray, batch_size: int, label_index: int, num_classes: int):
        if not (0 <= label_index < records.shape[1]):
            raise ValueError(f"labelIndex {label_index} outside the {records.shape[1]} columns")
        self.records, self.batch_size, self.label_index, self.num_classes = records, batch_size, label_index, num_classes
        lab = records[:, label_index]
        if np.any(lab < 0) or np.any(lab >= num_classes) or np.any(lab != np.round(lab)):
            raise ValueError("label column holds values outside [0, numClasses)")
        self._pos = 0

    def has_next(self) -> bool:
        return self._pos < len(self.records)

    def reset(self):
        self._pos = 0

    def next(self) -> Tuple[np.ndarray, np.ndarray]:
        if not self.has_next():
            raise StopIteration
        r = self.records[self._pos:self._pos + self.batch_size]
        self._pos += len(r)
        feats = np.ascontiguousarray(np.delete(r, self.label_index, axis=1), dtype=np.float32)
        labels = np.zeros((len(r), self.num_classes), np.float32)
        labels[np.arange(len(r)), r[:, self.label_index].astype(np.int64)] = 1.0
        return feats, labels

    def __iter__(self) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
        self.reset()
        while self.has_next():
            yield self.next()


def latent_grid(num_gen_samples: int = 10) -> np.ndarray:
    """J:382-389: the numGenSamples x numGenSamples grid of z in [-1,1]^2 (linspace on both axes, i outer, j inner)."""
    g = np.linspace(-1.0, 1.0, num_gen_samples, dtype=np.float32)
    return np.array([[g[i], g[j]] for i in range(num_gen_samples) for j in range(num_gen_samples)], np.float32)


def write_csv(path: str, features: np.ndarray, labels: np.ndarray):
    """The notebook's writer (np.savetxt(..., fmt="%.2f", delimiter=",") of [pixels | label], N:104-151) -- used by tests as the fixture maker."""
    np.savetxt(path, np.concatenate([features.reshape(len(features), -1), labels.reshape(-1, 1)], axis=1), fmt="%.2f", delimiter=",")
