// api_backward.inl — backward launch sequence + PPO loss (part of api.cu).
