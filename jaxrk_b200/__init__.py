"""jaxrk_b200 -- B200-native Gram / fused K@W path behind JaxRK's kernel / reduce / rkhs API."""
__version__ = "0.1.0"
