"""iswm_b200 — B200-native (sm_100a) implementation of the ISWM DeepLabV3+ hot path.

Public surface mirrors the reference's: `network.modeling.deeplabv3plus_resnet50`,
the class-weighted CE criterion built by train.py, `metrics.StreamMetrics`.
"""
__version__ = "0.1.0"
