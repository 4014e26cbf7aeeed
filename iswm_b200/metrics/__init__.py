from .stream_metrics import StreamMetrics

StreamSegMetrics = StreamMetrics      # upstream (VainF) name used by BASELINE.json's north_star

__all__ = ["StreamMetrics", "StreamSegMetrics"]
