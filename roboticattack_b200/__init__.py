"""B200-native adversarial-patch optimisation engine behind the roboticAttack ``OpenVLAAttacker`` API."""
__version__ = "0.1.0"
