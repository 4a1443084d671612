"""Dataset duck-type the SequentialVAE / trainer rely on (reference dataset/dataset.py:2-31) and a synthetic source of
the benchmark shapes (there is no network for MNIST / CelebA / LSUN)."""
import numpy as np


class Dataset:
    """Same attribute / method surface as the reference base class (dataset/dataset.py:2-31)."""

    def __init__(self):
        self.batch_size = -1
        self.name = "abstract"
        self.data_dims = []
        self.width = -1
        self.height = -1
        self.train_size = -1
        self.test_size = -1
        self.range = [0.0, 1.0]

    def next_batch(self, batch_size):
        self.handle_unsupported_op()

    def next_test_batch(self, batch_size):
        self.handle_unsupported_op()

    def display(self, image):
        return image

    def reset(self):
        self.handle_unsupported_op()

    def handle_unsupported_op(self):
        print("Unsupported Operation")
        raise Exception("Unsupported Operation")


_SHAPES = {
    # name -> (data_dims, range): dataset_mnist.py:13,21 (28x28 edge-padded to 32x32x1, [0,1]); dataset_cifar.py:9-10;
    # dataset_celeba.py:23,29 and dataset_lsun.py (64x64x3, [-1,1]); dataset_svhn.py (32x32x3, [0,1])
    "mnist": ([32, 32, 1], [0.0, 1.0]),
    "cifar": ([32, 32, 3], [0.0, 1.0]),
    "svhn": ([32, 32, 3], [0.0, 1.0]),
    "celebA": ([64, 64, 3], [-1.0, 1.0]),
    "lsun": ([64, 64, 3], [-1.0, 1.0]),
}


class SyntheticDataset(Dataset):
    """x ~ U[range] NHWC float32 batches of a named dataset's shape, seeded (SURVEY.md 8d 'Synthetic inputs')."""

    def __init__(self, name="celebA", batch_size=100, seed=1234, data_dims=None, data_range=None):
        super().__init__()
        dims, rng = _SHAPES.get(name, (data_dims, data_range))
        self.name = name
        self.data_dims = list(data_dims if data_dims is not None else dims)
        self.range = list(data_range if data_range is not None else rng)
        self.height, self.width = self.data_dims[0], self.data_dims[1]
        self.batch_size = batch_size
        self.train_size = self.test_size = 1 << 20
        self._seed = seed
        self.reset()

    def _draw(self, rng, n):
        lo, hi = self.range
        return rng.uniform(lo, hi, size=[n] + self.data_dims).astype(np.float32)

    def next_batch(self, batch_size):
        return self._draw(self._train_rng, batch_size)

    def next_test_batch(self, batch_size):
        return self._draw(self._test_rng, batch_size)

    def reset(self):
        self._train_rng = np.random.default_rng(self._seed)
        self._test_rng = np.random.default_rng(self._seed + 1)
