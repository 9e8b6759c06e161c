"""print_time and the loss callbacks of lavis/compression/pruners/utils.py (:6-44) the path keeps."""
import functools
import time


def print_time(func):
    @functools.wraps(func)
    def timed(*args, **kwargs):
        t0 = time.time()
        out = func(*args, **kwargs)
        print(f"{func.__qualname__} took {time.time() - t0:.4f} s")
        return out
    return timed


def prepare_sample(samples, cuda_enabled=True):
    """lavis.datasets.data_utils.prepare_sample as the pruners use it (utils.py:22): tensors of the batch to the GPU."""
    if not cuda_enabled:
        return samples
    import torch

    def move(x):
        if torch.is_tensor(x):
            return x.cuda(non_blocking=True)
        if isinstance(x, dict):
            return {k: move(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(move(v) for v in x)
        return x
    return move(samples)


def loss_vision_language(model, samples, cuda_enabled):
    """utils.py:21-31 (loss_language, :34-44, is the same function): (loss, batch length) of one calibration batch."""
    samples = prepare_sample(samples, cuda_enabled=cuda_enabled)
    loss = model(samples)["loss"]
    return loss, len(samples["text_input"])


loss_language = loss_vision_language
