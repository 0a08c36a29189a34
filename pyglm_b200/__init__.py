"""pyglm_b200: B200-native Gibbs sampler behind PyGLM's sparse Bernoulli network GLM API."""
__version__ = "0.1.0"
