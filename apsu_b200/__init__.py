"""apsu_b200 — B200-native receiver-side homomorphic query evaluation for APSU.

Only the hot path of SURVEY.md §8: ComputePowers + BatchedPlaintextPolyn::eval/eval_patstock behind
the reference's Receiver / ReceiverDB / BinBundle / PSUParams surface.  All arithmetic runs in
libapsu_b200.so (hand-written sm_100a CUDA, apsu_b200/csrc); there is no CPU fallback.
"""
from .capi import CudaUnavailable, LIB_PATH  # noqa: F401
from .receiver import PSUParams, PowersDag, PowersNode, Query, Receiver, ReceiverDB, ResultPackage  # noqa: F401
