"""Import stub so the reference's palette.py loads offline (test infrastructure only).

colormath 3.0.0 is a requirements.txt dependency of the reference that is not
installed here; palette.py only constructs sRGBColor(r, g, b, is_upscaled=True)
objects (reference transcoder/palette.py:6-15).
"""
