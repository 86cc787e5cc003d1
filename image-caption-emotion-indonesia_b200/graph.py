"""CUDA-graph capture of one whole training step (forward + fused loss + backward + clamp/Adam, and the
gradient all-reduce when data parallel).  The step issues ~40 kernels whose total device time at the reference
batch sizes (64-96) is ~1 ms: launched one by one from Python the step is launch-bound, replayed from a graph
it is not.  Everything step-dependent lives in device memory (dropout counter, Adam step counters, learning
rate), so a replay is a real, fresh training step.

Only fully teacher-forced steps are captured (teacher_forcing_ratio >= 1): scheduled sampling draws a host
coin per time step (stylenet/model.py:181) that changes the kernel sequence."""
import torch


class GraphedTrainStep:
    def __init__(self, trainer, captions, lengths, features, warmup=3, **step_kw):
        """``trainer``: DataParallelTrainer (world size 1 or more).  ``captions`` / ``features``: example device
        tensors defining the static input shapes; ``lengths`` is fixed for the life of the graph."""
        if step_kw.get("teacher_forcing_ratio", 1.0) < 1.0:
            raise ValueError("only teacher-forced steps can be captured")
        self.trainer = trainer
        self.lengths = list(lengths)
        self.kw = step_kw
        self.captions = captions.clone()
        self.features = features.clone() if features is not None else None
        dec = trainer.decoder
        # warm up on a side stream (allocator, plans, lazily built state), then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.stats = self._run()
        self.n_kernels = None
        _ = dec

    def _run(self):
        return self.trainer.step(self.captions, self.lengths, self.features, **self.kw)

    def __call__(self, captions=None, features=None):
        """Copy new inputs into the static buffers (async, same stream) and replay.  Returns the static loss
        tensor (valid after the stream reaches this point)."""
        if captions is not None:
            self.captions.copy_(captions, non_blocking=True)
        if features is not None:
            self.features.copy_(features, non_blocking=True)
        self.graph.replay()
        return self.loss, self.stats
