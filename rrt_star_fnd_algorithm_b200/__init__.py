"""B200-native hot path of the RRT / RRT* / RRT*-FN / RRT*-FND planners."""
