"""`hybrid_models.model_hybrid` as the reference's drivers import it -> the B200-native implementation."""
from estdepth_b200.model import DepthNetHybrid  # noqa: F401
