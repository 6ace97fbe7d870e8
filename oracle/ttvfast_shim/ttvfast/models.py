"""ttvfast.models.Planet as used by nauyaca/utils.py:60-70."""


class Planet:
    def __init__(self, mass, period, eccentricity, inclination, longnode, argument, mean_anomaly):
        self.mass, self.period, self.eccentricity = mass, period, eccentricity
        self.inclination, self.longnode, self.argument, self.mean_anomaly = inclination, longnode, argument, mean_anomaly
