"""async-ev-cnn on B200: the event-driven EFCN inference hot path as hand-written sm_100a CUDA
behind a C ABI, with a Python host layer that mirrors the reference's Layer / model-builder /
run_networks surface.  See DESIGN.md.  (Import as `async_ev_cnn_b200`.)"""
from .streams import EFCN_LAYERS, parse_layers, synthetic_events, xavier_weights  # noqa: F401
