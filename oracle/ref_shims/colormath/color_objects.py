"""Minimal stand-ins for the two colormath classes the reference names."""


class sRGBColor:
    def __init__(self, rgb_r, rgb_g, rgb_b, is_upscaled=False):
        self.is_upscaled = is_upscaled
        self.rgb_r, self.rgb_g, self.rgb_b = rgb_r, rgb_g, rgb_b

    def get_upscaled_value_tuple(self):
        return (int(self.rgb_r), int(self.rgb_g), int(self.rgb_b))


class LabColor:
    def __init__(self, lab_l, lab_a, lab_b):
        self.lab_l, self.lab_a, self.lab_b = lab_l, lab_a, lab_b
