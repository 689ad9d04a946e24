"""Data parallelism as the reference does it: every rank holds a
full replica, the batch is split contiguously over ranks (run.py:287-296), and the
only collective is a SUM all-reduce of gradients followed by x 1/num_ranks
(tfutil.py:326-344).  The reference issues one nccl.all_sum per variable (96 for
the E+G optimizer); here the gradients of ALL networks of an optimizer phase live in
ONE flat fp32 bucket (optim.GradientBucket, non-finite marks in its tail), so a train
step is two NCCL all-reduces over NVLink 5 / NVSwitch.  One process per GPU (torchrun /
torch.distributed)."""
import os

import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_initialized() else 1


def rank():
    return dist.get_rank() if is_initialized() else 0


def init_from_env(backend=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).
    NCCL when CUDA is present, gloo otherwise (CPU tests of the host logic)."""
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if ws <= 1 or is_initialized():
        return
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    kwargs = {}
    if backend == 'nccl':
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        kwargs['device_id'] = torch.device('cuda', local)
    dist.init_process_group(backend, **kwargs)


def shard_bounds(num_items, num_ranks=None, r=None):
    """Contiguous split of a global batch like tf.split(reals, num_gpus) (run.py:287): [begin, end) of rank r."""
    num_ranks = world_size() if num_ranks is None else num_ranks
    r = rank() if r is None else r
    if num_items % num_ranks != 0:
        raise ValueError('global batch %d is not divisible by %d ranks (run.py:287 tf.split)' % (num_items, num_ranks))
    per = num_items // num_ranks
    return r * per, (r + 1) * per


def allreduce_sum_(flat_buffers):
    """In-place SUM all-reduce of each flat gradient buffer (one collective per buffer).
    No-op in a single-process job.  Returns the number of collectives issued."""
    if world_size() == 1:
        return 0
    n = 0
    for buf in flat_buffers:
        if buf.numel() == 0:          # nccl does not support zero-sized tensors (tfutil.py:329)
            continue
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        n += 1
    return n


def allreduce_sum_async_(flat_buffer):
    """Issue the in-place SUM all-reduce of one flat buffer WITHOUT making the current stream wait for it: the
    collective starts once the kernels enqueued so far have finished and runs on the communicator's own stream, so
    kernels launched next overlap it.  Returns wait() - call it before anything reads the buffer (it makes the current
    stream wait, the host does not block).  Single-process job: nothing is issued, wait() is a no-op."""
    if world_size() == 1 or flat_buffer.numel() == 0:
        return lambda: None
    work = dist.all_reduce(flat_buffer, op=dist.ReduceOp.SUM, async_op=True)
    return work.wait


def broadcast_(flat_buffers, src=0):
    """Make replicas bit-identical at start-up (the reference clones variables per tower, run.py:303-309)."""
    if world_size() == 1:
        return
    for buf in flat_buffers:
        dist.broadcast(buf, src=src)
