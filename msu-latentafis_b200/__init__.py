"""B200-native 1-vs-N latent fingerprint gallery matcher (MSU-LatentAFIS `matching/` drop-in)."""
