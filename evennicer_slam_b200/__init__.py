"""B200-native fused ray renderer for EvenNICER-SLAM (hot path only).

Nothing is imported here: ``import evennicer_slam_b200`` must work without torch or the CUDA library (the numpy-only
``synthetic`` module is used by the oracle tests).  Import what you need:

    renderer.Renderer, decoder.NICE, common.get_samples / get_camera_from_tensor     drop-ins for src/utils/Renderer.py,
                                                                                      src/conv_onet/models/decoder.py, src/common.py
    event_net.inference_event, mapper_ops.FrustumSelector                             drop-ins for src/event_net.py, Mapper methods
    optim.FrustumGridAdam / FusedAdam, losses.mapper_loss / tracker_loss / event_loss optional fused caller-side ops
    graph.GraphedStep, sharding.*                                                     CUDA-graph replay, one-process-per-GPU sharding
    _lib (ctypes binding of include/ens_render.h), _ext (torch C++ extension)         the two bindings of libens_render.so
"""
__version__ = "0.2.0"
__all__ = ["renderer", "decoder", "common", "event_net", "mapper_ops", "optim", "losses", "graph", "sharding", "scene",
           "functional", "harness", "synthetic"]
