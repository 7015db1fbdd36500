"""molgym_b200 — B200-native PPO policy/value hot path of gncs/molgym behind the reference's agent API."""
__version__ = '0.1.0'
