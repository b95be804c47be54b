"""Parity oracle (TEST INFRASTRUCTURE ONLY): see oracle/hyperion_oracle.cpp."""
