"""Drop-in for lidbox/models/xvector_extended.py (Villalba et al. 2018 extended x-vector) on the same sm_100a kernels.

    m = create(input_shape=(T or None, F), num_outputs, output_activation="log_softmax")   # xvector_extended.py:22-43
    emb = as_embedding_extractor(m)(x)                                                     # re-exported, :18

Ten causal frame layers (kernel/stride 5/1, 1/1, 3/2, 1/1, 3/3, 1/1, 3/4, 1/1, 1/1, 1/1), statistics pooling, two
segment layers and the `output` Dense.  Every layer is the implicit-GEMM frame layer of models/xvector.py; the
stride-4 layer has kernel_size < strides, i.e. its TMA view has gaps instead of overlaps and the skipped input frames
receive a zero data gradient.
"""
from .xvector import (                                           # noqa: F401  (same re-exports as the reference)
    XVector,
    frame_layer,
    GlobalMeanStddevPooling1D,
    segment_layer,
    as_embedding_extractor,
)

_HEADS = {"log_softmax": "log_softmax", None: "none", "": "none"}


def frame_layers():
    """xvector_extended.py:25-34."""
    return [frame_layer(512, 5, 1, name="frame1"), frame_layer(512, 1, 1, name="frame2"),
            frame_layer(512, 3, 2, name="frame3"), frame_layer(512, 1, 1, name="frame4"),
            frame_layer(512, 3, 3, name="frame5"), frame_layer(512, 1, 1, name="frame6"),
            frame_layer(512, 3, 4, name="frame7"), frame_layer(512, 1, 1, name="frame8"),
            frame_layer(512, 1, 1, name="frame9"), frame_layer(1500, 1, 1, name="frame10")]


def create(input_shape, num_outputs, output_activation="log_softmax", **kwargs):
    """xvector_extended.py:22-43.  output_activation: "log_softmax" (default) or None (raw scores of `output`)."""
    if output_activation not in _HEADS:
        raise NotImplementedError("output_activation %r: only 'log_softmax' and None are implemented"
                                  % (output_activation,))
    return XVector(input_shape, num_outputs, name="x-vector-extended", frames=frame_layers(),
                   head=_HEADS[output_activation], output_name="output", **kwargs)
