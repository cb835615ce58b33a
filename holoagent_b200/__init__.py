"""holoagent_b200: B200-native (sm_100a) HMSG build-and-retrieve hot path of
HorizonRobotics/HoloAgent's FSR-VLN behind the reference's perception / memory API.

    from holoagent_b200.engine import HmsgEngine            # C-ABI owner (one per GPU)
    from holoagent_b200.memory.hmsg.graph.graph import Graph # drop-in for the reference Graph hot path

There is no CPU fallback: the CUDA library (libhmsg_b200.so) must be built and a B200 present.
"""
__version__ = "0.1.0"
