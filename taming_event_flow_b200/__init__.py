"""B200-native (sm_100a) contrast-maximization hot path of tudelft/taming_event_flow.

The Python surface mirrors the reference so that its scripts can switch imports:

    from taming_event_flow_b200.utils.iwe import get_interpolation, interpolate, ...
    from taming_event_flow_b200.loss.flow import Linear, Iterative
    from taming_event_flow_b200.dataloader.encodings import events_to_voxel, events_to_channels

Everything ends in hand-written CUDA kernels behind the C ABI of ``include/tef_b200.h``.
"""
__version__ = "0.1.0"
