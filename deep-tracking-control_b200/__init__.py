"""dtc_b200: B200-native hot path of priest-yang/Deep-Tracking-Control."""
__version__ = "0.1.0"
