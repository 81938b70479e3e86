"""Drop-in replacements for the reference's ``networks`` package (same import paths, class names,
constructor signatures and state_dict keys); the compute runs on libustrun_sm100.so."""
