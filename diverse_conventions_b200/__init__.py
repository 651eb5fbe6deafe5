"""B200-native batched Overcooked / Balance-Beam simulator and rollout engine.

Drop-in for the hot path of Stanford-ILIAD/Diverse-Conventions behind the
``pantheonrl_extension.VectorMultiAgentEnv`` API.  See DESIGN.md.
"""
__version__ = "0.1.0"
