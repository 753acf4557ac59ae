"""pywgcna_b200 -- B200-native (sm_100a) network construction for PyWGCNA.

Drop-in for the reference's hot path only: ``pickSoftThreshold``, ``adjacency``, ``TOMsimilarity``
and the connectivity helpers (``PyWGCNA/wgcna.py:788-1193, 3342-3450``), behind the reference's own
signatures.  ``install()`` rebinds them on ``PyWGCNA.wgcna.WGCNA``.
"""
from .wgcna import (TOMDenoms, TOMTypes, TOMsimilarity, TomSimilarityFromAdj, adjacency, adjacencyTypes,
                    calBlockSize, checkAdjMat, checkSimilarity, connectivity, dissTOM, install, intramodularConnectivity, linkage, network,
                    network_blocks, networkTypes, pickSoftThreshold, scaleFreeFitIndex, set_correlation_precision, signedKME, softConnectivity, uninstall)

__version__ = "0.1.0"
__all__ = ["TOMDenoms", "TOMTypes", "TOMsimilarity", "TomSimilarityFromAdj", "adjacency", "adjacencyTypes",
           "calBlockSize", "checkAdjMat", "checkSimilarity", "connectivity", "dissTOM", "install", "intramodularConnectivity",
           "linkage", "network", "network_blocks", "networkTypes", "pickSoftThreshold", "scaleFreeFitIndex", "set_correlation_precision", "signedKME", "softConnectivity", "uninstall"]
