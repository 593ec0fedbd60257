"""B200-native NTN cost+gradient hot path behind the reference's NTN / DataFactory API.

The compute path is the C-ABI library ``libntn_b200.so`` (hand-written CUDA for
sm_100a, `csrc/`); this package is the host-side mirror of the reference's
interface for that path (/root/reference/src/NTN.java, DataFactory.java,
Run_NTN.java).  There is no CPU fallback: importing the compute classes without
the built library raises.
"""
__all__ = ["synth"]
