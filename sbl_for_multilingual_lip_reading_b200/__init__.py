"""B200-native visual encoder (Conv3d frontend + ResNet-18 trunk + transformer Encoder) for
SBL_For_Multilingual_Lip_Reading — drop-in nn.Module replacements backed by libsblk.so (sm_100a)."""
__version__ = "0.1.0"
