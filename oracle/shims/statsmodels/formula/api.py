ols = gls = wls = glm = None
