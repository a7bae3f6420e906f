"""B200-native training hot path of yyNoBug/VAE_segmentation (drop-in modules + losses)."""
__version__ = "0.1.0"
