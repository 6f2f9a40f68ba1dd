AxesGrid = None
