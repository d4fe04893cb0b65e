"""Empty stand-in: the reference tree graphics import matplotlib at module load (oracle test infra only)."""
