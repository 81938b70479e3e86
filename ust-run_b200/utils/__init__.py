"""Drop-in replacements for the hot-path parts of the reference's ``utils`` package."""
