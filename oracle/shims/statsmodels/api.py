OLS = GLS = WLS = None
