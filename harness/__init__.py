"""Test / bench harness (scene generators, oracle loader). Not part of the product path."""
