"""ml-100k example workflow on the B200 path (mirror of the reference's recman/examples; BASELINE config 1)."""
