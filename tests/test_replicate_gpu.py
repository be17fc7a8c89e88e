"""Single-GPU checks of the operand-replication kernels (include/b200_replicate.h): data copy to one
or several ordinary destinations, arrival flag published after the data, wait / signal kernels.
The multicast and peer-mapped forms need several GPUs: tools/multi_gpu_check.py --bcast nvlink."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbytes", [16, 4096, 1 << 20, (1 << 24) + 48])
@pytest.mark.parametrize("n_dst", [1, 3])
def test_push_copies_and_publishes_flag(ob, nbytes, n_dst):
    import torch
    n = nbytes // 4
    g = torch.Generator(device="cuda").manual_seed(nbytes + n_dst)
    src = torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
    dsts = [torch.zeros(n, device="cuda", dtype=torch.int32) for _ in range(n_dst)]
    flags = torch.zeros(16, device="cuda", dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    before = ob.launch_count()
    ob.replicate_push([d.data_ptr() for d in dsts], src.data_ptr(), nbytes,
                      [flags.data_ptr() + 4 * i for i in range(n_dst)], 7, multicast=False, flag_multicast=False,
                      ctas=5, stream=st)
    ob.flag_wait(flags.data_ptr(), 7, count=n_dst, stride=1, skip=-1, stream=st)
    ob.flag_signal(flags.data_ptr() + 4 * 9, 0xFFFFFFFF, stream=st)
    torch.cuda.synchronize()
    assert ob.launch_count() - before == 3
    for d in dsts:
        assert torch.equal(d, src)
    f = flags.cpu().numpy().view(np.uint32)
    assert list(f[:n_dst]) == [7] * n_dst and f[9] == 0xFFFFFFFF and f[n_dst:9].sum() == 0


def test_flag_wait_is_wrap_safe_and_skips(ob):
    import torch
    flags = torch.tensor([5, 0, 0xFFFFFFF0 - 2**32, 6], device="cuda", dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    ob.flag_wait(flags.data_ptr(), 5, count=2, stride=3, skip=-1, stream=st)     # lanes watch words 0 and 3 (5 and 6)
    ob.flag_wait(flags.data_ptr(), 5, count=2, stride=1, skip=1, stream=st)      # word 1 (0: would never arrive) is skipped
    ob.flag_wait(flags.data_ptr() + 8, 0xFFFFFFE0, count=1, stream=st)           # sequence numbers near the wrap: F0 has reached E0
    torch.cuda.synchronize()


def test_push_rejects_misaligned(ob):
    import torch
    src = torch.zeros(64, device="cuda", dtype=torch.int32)
    dst = torch.zeros(64, device="cuda", dtype=torch.int32)
    with pytest.raises(ob.B200Error):
        ob.replicate_push([dst.data_ptr()], src.data_ptr(), 24, [], 0, multicast=False, flag_multicast=False)
    with pytest.raises(ob.B200Error):
        ob.replicate_push([dst.data_ptr() + 4], src.data_ptr(), 32, [], 0, multicast=False, flag_multicast=False)
