"""Drop-in shim: lets the reference's drivers (`from hybrid_models.model_hybrid import DepthNetHybrid`,
eval_hybrid.py:11, eval_hybrid_seq.py:10) import the B200-native model unchanged when this repository precedes the
reference tree on sys.path."""
