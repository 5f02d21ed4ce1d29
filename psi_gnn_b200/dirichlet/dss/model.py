"""Drop-in for the reference's ``dirichlet/dss/model.py`` (inference on the shared fused layer kernel)."""
from ...baselines import DeepStatisticalSolver, Psi, MLPActivation                       # noqa: F401
from ...baselines import DecoderDSS as Decoder                                           # noqa: F401
from ...model import MLP, Phi_to, Phi_from, initialize_weights_xavier                    # noqa: F401
