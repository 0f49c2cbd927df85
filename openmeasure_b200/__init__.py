"""openmeasure_b200 -- B200-native snapshot-POD sparse-sensing hot path (ROM / SPR).

    from openmeasure_b200.sparse_sensing import ROM, SPR

mirrors `openmeasure.sparse_sensing` of the reference for the path fit -> optimal_placement ->
train -> predict -> reconstruct.  The CUDA library is built in-tree by `openmeasure_b200.build`.
"""
__version__ = "0.1.0"
