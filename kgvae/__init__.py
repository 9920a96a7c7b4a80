"""Drop-in `kgvae` package: the reference's module / CLI interface for the KG-VAE training hot path,
re-hosted on ark_b200's sm_100a kernels."""
