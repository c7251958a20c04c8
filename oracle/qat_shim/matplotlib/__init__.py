"""TEST INFRASTRUCTURE ONLY -- import stand-in for matplotlib (absent in the build image): the reference's facade
(openvqe/algorithms/algorithm.py:1) imports pyplot at module level; nothing here draws anything."""
