"""Minimal stand-in for Biopython: exactly the surface Mgcod's bin/*.py touch."""
