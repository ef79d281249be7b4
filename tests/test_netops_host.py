"""Host-side logic of the deferred weight gradient (netops.WindowStacks) and of the fused operators' applicability tests; CPU only
(the kernels themselves are covered on the GPU by tests/test_netops_gpu.py)."""
import torch

from taming_event_flow_b200 import netops


def test_window_stacks_hand_out_one_slot_per_forward_call_of_an_open_window():
    st = netops.WindowStacks()
    assert st.take_slot() is None                     # no window open: per-pass gradients
    st.begin(3)
    assert [st.take_slot() for _ in range(4)] == [0, 1, 2, None]      # a fourth call falls back
    assert st.used == 3
    st.begin(2)
    assert st.used == 0 and st.take_slot() == 0
    with torch.no_grad():
        assert st.take_slot() is None                 # no gradients wanted: nothing to defer
    assert st.take_slot() == 1 and st.used == 2
    st.begin(0)
    assert st.take_slot() is None


def test_window_stacks_allocate_channels_last_rows_per_pass():
    st = netops.WindowStacks()
    st.begin(4)
    w = torch.zeros(8, 6, 3, 3)
    x0 = netops.conv_input_slot(st, 0, (2, 6, 10, 12), w, 2, 1, torch.device("cpu"))
    x3 = netops.conv_input_slot(st, 3, (2, 6, 10, 12), w, 2, 1, torch.device("cpu"))
    assert tuple(st.x["x"].shape) == (8, 6, 10, 12) and tuple(st.g["y"].shape) == (8, 8, 5, 6)      # stride 2, padding 1
    assert x0.is_contiguous(memory_format=torch.channels_last) and tuple(x0.shape) == (2, 6, 10, 12)
    assert x3.data_ptr() == st.x["x"].data_ptr() + 3 * 2 * 6 * 10 * 12 * 4
    buf = st.x["x"]
    netops.conv_input_slot(st, 1, (2, 6, 10, 12), w, 2, 1, torch.device("cpu"))
    assert st.x["x"] is buf                           # same geometry: no re-allocation
    netops.conv_input_slot(st, 1, (2, 6, 12, 12), w, 2, 1, torch.device("cpu"))
    assert st.x["x"] is not buf and tuple(st.x["x"].shape) == (8, 6, 12, 12)


def test_fused_operators_only_claim_tensors_they_can_take():
    x = torch.zeros(1, 8, 4, 4)
    assert not netops.usable(x) and not netops.usable_input(x)          # CPU tensors: the modules run their plain formulation
    assert netops.usable(None)
    assert netops.ACT == {None: 0, "none": 0, "relu": 1, "tanh": 2}
