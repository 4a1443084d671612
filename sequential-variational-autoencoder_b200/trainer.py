"""NoisyTrainer - the caller either side of the hot path (reference trainer.py:23-206), kept so that ``main.py``-style
drivers run unchanged: ``NoisyTrainer(network, dataset, args, logger, base_dir).train()``.

What differs from the reference: the denoising corruption (``apply_noise``, trainer.py:56-78) runs on the device
(``svae_apply_noise`` / ``svae_train_step_host_denoise``: Philox draws, one upload of the clean batch instead of input +
target), and the matplotlib / scipy.misc plotting (``plot_reconstruction``, trainer.py:143-206) is not part of this path.
"""
import time

import numpy as np


class NoisyTrainer:
    # trainer.py:16-21
    pepper_prob = 0.1
    salt_prob = 0.1
    gaussian_noise_scale = 0.1
    max_training_iters = 10000000
    test_num_iters = 5
    log_loss_freq = 20

    def __init__(self, network, dataset, args, logger, base_dir):
        self.network = network
        self.dataset = dataset
        self.args = args
        self.batch_size = args.batch_size
        self.data_dims = self.dataset.data_dims
        self.fig = None
        self.LOG = logger
        self.base_dir = base_dir
        self._noise_calls = 0

    def apply_noise(self, original):
        """trainer.py:56-78 on the device; raises like the reference when denoise_train is off (:66-67)."""
        if not self.args.denoise_train:
            raise Exception("Called apply_noise, but self.args.denoise_train==False, is this right?")
        self._noise_calls += 1
        return self.network.apply_noise(original, NoisyTrainer.pepper_prob, NoisyTrainer.salt_prob,
                                        NoisyTrainer.gaussian_noise_scale, seed=(1 << 32) + self._noise_calls)

    def train(self, max_iters=None):
        """trainer.py:82-110: test + visualise every vis_frequency iterations, one network.train per iteration, loss logged
        every log_loss_freq iterations.  Returns the last training loss."""
        train_loss = float("nan")
        n = NoisyTrainer.max_training_iters if max_iters is None else int(max_iters)
        for iteration in range(n):
            iter_beg_time = time.time()
            if iteration % self.args.vis_frequency == 0:
                test_error = self.test(iteration // self.args.vis_frequency)
                self.LOG.info("Reconstruction error per pixel: %f, @ iteration: %d" % (test_error, iteration))
                self.network.visualize(iteration // self.args.vis_frequency)
            target_batch = self.dataset.next_batch(self.batch_size)
            if self.args.denoise_train:
                train_loss = self.network.train_denoise(target_batch, NoisyTrainer.pepper_prob, NoisyTrainer.salt_prob,
                                                        NoisyTrainer.gaussian_noise_scale)
            else:
                train_loss = self.network.train(target_batch, target_batch)
            if iteration % NoisyTrainer.log_loss_freq == 0:
                self.LOG.info("Iteration %d: Reconstruction loss %f, time per iter %fs" %
                              (iteration, train_loss, time.time() - iter_beg_time))
        return train_loss

    def test(self, epoch, num_iters=None):
        """trainer.py:114-141: mean over num_iters test batches of sum((reconstruction - target)^2) / (H*W) / batch."""
        num_iters = NoisyTrainer.test_num_iters if num_iters is None else num_iters
        error = 0.0
        for _ in range(num_iters):
            test_target_batch = self.dataset.next_test_batch(self.batch_size)
            test_input_batch = self.apply_noise(test_target_batch) if self.args.denoise_train else test_target_batch
            reconstruction = self.network.test(test_input_batch)
            error += np.sum(np.square(reconstruction - test_target_batch)) / np.prod(self.data_dims[:2]) / self.batch_size
        return error / num_iters
