"""Drop-in modules with the file / class / parameter names of the reference's `layers/` package."""
