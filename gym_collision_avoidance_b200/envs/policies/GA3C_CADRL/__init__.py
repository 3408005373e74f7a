"""GA3C-CADRL network (reference ``envs/policies/GA3C_CADRL``): TensorFlow-free loader + device inference."""
