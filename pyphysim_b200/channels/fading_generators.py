"""placeholder filled in below"""
