"""q1tsim_b200 -- B200-native statevector engine behind q1tsim's VectorState."""
__version__ = "0.1.0"
