"""Import-only stand-in: the reference imports matplotlib in the renderer/generator modules but never plots on the hot path."""
