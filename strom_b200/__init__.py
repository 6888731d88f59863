"""strom_b200 -- B200-native (sm_100a) tree-likelihood engine behind strom's BeagleLib boundary."""
__version__ = "0.1.0"
