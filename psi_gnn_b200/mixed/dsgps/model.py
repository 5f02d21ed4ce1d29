"""Drop-in for the reference's ``mixed/dsgps/model.py`` (GRU-gated recurrent baseline with Dirichlet + Neumann boundary conditions)."""
from ...baselines import ModelDSGPSMixed as ModelDSGPS, MLPActivation, Psi      # noqa: F401
from ...model import MLP, Phi_to, Phi_from, Encoder, Decoder, Autoencoder        # noqa: F401
